"""kimimaro_b200/csrc/preamble.cu on the CPU: the kernels (connected components, hole filling) are compiled by g++
against the SIMT emulation (tests/host/emu_include/cuda_runtime.h; no block-level synchronisation in this file, so a
launch is a loop over blocks and threads) and the library's own entry points run on host arrays against the oracle.
The hole-filling kernel then stands in for itself in the host logic of fill_holes / fix_avocados
(kimimaro_b200/intake.py), including the three-slice sandwich that turns it into the 2-D fill of paint_walls."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "host", "preamble_emu.cpp")
DEPS = [SRC, os.path.join(HERE, "host", "emu_include", "cuda_runtime.h"), os.path.join(HERE, "host", "emu_include", "simt_impl.h"),
        os.path.join(ROOT, "kimimaro_b200", "csrc", "preamble.cu"), os.path.join(ROOT, "kimimaro_b200", "csrc", "common.cuh")]
OUT = os.path.join(ROOT, "oracle", "_cache", "preamble_emu.so")
c_i64, c_u64, c_vp = ctypes.c_int64, ctypes.c_uint64, ctypes.c_void_p


@pytest.fixture(scope="module")
def emu():
  os.makedirs(os.path.dirname(OUT), exist_ok=True)
  if (not os.path.exists(OUT)) or os.path.getmtime(OUT) < max(os.path.getmtime(d) for d in DEPS):
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wno-attributes",
                           "-I" + os.path.join(HERE, "host", "emu_include"), "-I" + os.path.join(HERE, "host"), SRC, "-o", OUT])
  return ctypes.CDLL(OUT)


def _ccl(lib, lab):
  """engine.connected_components with host arrays."""
  sx, sy, sz = lab.shape
  flat = np.ascontiguousarray(lab.reshape(-1, order="F"))
  parent = np.empty(flat.size, np.uint32)
  is_root = np.empty(flat.size, np.uint8)
  assert lib.b2t_ccl26_roots(oracle._p(flat), flat.dtype.itemsize, c_i64(sx), c_i64(sy), c_i64(sz), oracle._p(parent),
                             oracle._p(is_root), None) == 0
  rank = np.cumsum(is_root, dtype=np.int32)
  assert lib.b2t_ccl_relabel(oracle._p(parent), oracle._p(rank), c_u64(flat.size), None) == 0
  return parent.reshape(lab.shape, order="F"), int(rank[-1])


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.uint32, np.uint64])
def test_ccl_kernels_equal_oracle(emu, dtype):
  rng = np.random.default_rng(21)
  for trial in range(8):
    shape = tuple(int(v) for v in rng.integers(1, 28, size=3))
    lab = rng.integers(0, 4, size=shape).astype(dtype)
    rep = tuple(int(v) for v in rng.integers(1, 4, size=3))
    lab = np.asfortranarray(np.repeat(np.repeat(np.repeat(lab, rep[0], 0), rep[1], 1), rep[2], 2)[:shape[0], :shape[1], :shape[2]])
    got, n = _ccl(emu, lab)
    ref, n_ref = oracle.connected_components(lab)
    assert n == n_ref and np.array_equal(got, ref), (shape, trial)


def _kernel_fill(lib):
  """kimimaro_b200.intake._fill_voids with the emulated b2t_fill_voids behind it (torch CPU tensors)."""
  import torch

  def fill_fn(mask, cshape):
    m = np.ascontiguousarray(mask.numpy())
    V = m.size
    reach, queue, ctrl = np.empty(V, np.uint32), np.empty(2 * V, np.uint32), np.zeros(16, np.uint32)
    rc = lib.b2t_fill_voids(oracle._p(m), c_i64(cshape[0]), c_i64(cshape[1]), c_i64(cshape[2]), oracle._p(reach),
                            oracle._p(queue), c_u64(V), oracle._p(ctrl), None)
    assert rc == 0
    mask.copy_(torch.from_numpy(m))
    return int(ctrl[5])
  return fill_fn


def test_fill_kernels_equal_oracle_and_the_2d_sandwich(emu):
  import scipy.ndimage as ndi
  import torch
  from kimimaro_b200 import intake
  rng = np.random.default_rng(22)
  fill_fn = _kernel_fill(emu)
  for trial in range(12):
    shape = tuple(int(v) for v in rng.integers(3, 26, size=3))
    m = (rng.random(shape) < 0.55)
    m = np.asfortranarray(ndi.binary_closing(m, iterations=1) | m)
    ref = m.copy(order="F")
    _, k_ref = oracle.fill_voids(ref)
    t = torch.from_numpy(m.reshape(-1, order="F").astype(np.uint8))
    k = fill_fn(t, shape)
    assert k == k_ref and np.array_equal(t.numpy().reshape(shape, order="F").astype(bool), ref), (shape, trial)
    # paint_walls' 2-D fill (intake.py:666-677) through the 3-D kernel: the image between two solid slices
    plane = torch.from_numpy(np.ascontiguousarray(m[:, :, 0].T))           # [y, x], like a face of a [z, y, x] crop
    got2 = intake._fill_voids_2d(plane, fill_fn).numpy().T
    assert np.array_equal(got2, ndi.binary_fill_holes(m[:, :, 0])), (shape, trial)


def test_fill_holes_and_avocados_host_logic_on_the_real_fill_kernel(emu):
  """fill_all_holes and engage_avocado_protection of kimimaro_b200/intake.py with the library's own hole-filling
  kernels (emulated) instead of the oracle's fill: what the GPU runs for fill_holes=True / fix_avocados=True except for
  the EDT and the label statistics."""
  import tests.test_oracle_cpu as T
  import torch
  from kimimaro_b200 import intake
  from oracle import teasar
  import scipy.ndimage as ndi
  fill_fn = _kernel_fill(emu)
  v = T._holey_volume()
  cc, n = oracle.connected_components(v)
  ref = teasar.fill_all_holes(cc.copy(order="F"), n)
  d_cc = torch.from_numpy(cc.reshape(-1, order="F").astype(np.int32))
  count = np.bincount(cc.ravel(), minlength=n + 1)
  bbox = np.zeros((n + 1, 6), np.int32)
  for l, slc in enumerate(ndi.find_objects(cc, max_label=n), start=1):
    bbox[l] = [slc[0].start, slc[1].start, slc[2].start, slc[0].stop - 1, slc[1].stop - 1, slc[2].stop - 1]
  out = intake.fill_all_holes(d_cc, cc.shape, n, count, bbox, fill_fn=fill_fn)
  assert np.array_equal(out.numpy().reshape(cc.shape, order="F"), ref.astype(np.int32))
  # the avocado volume of automated_test.py:478-509 at half size, through the shared helper with the kernel's fill
  labels = np.zeros((128, 128, 128), dtype=np.uint32, order="F")
  labels[:25, :20, :15] = 1
  labels[:12, :10, :12] = 2
  labels[25:50, 20:50, 15:40] = 3
  labels[30:45, 25:45, 20:35] = 4
  labels[30:35, 26:44, 21:34] = 5
  labels[100:, 100:, 100:] = 6
  labels[75:100, 100:, 100:] = 7
  got, _ = T._avocado_host_vs_oracle(labels, 1, fill_fn=fill_fn)
  assert len(np.unique(got)) == 5
