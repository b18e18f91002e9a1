"""K1 column pass v3 on the CPU: tests/host/fh3_host.cpp instantiates the very template the CUDA kernel
instantiates (kimimaro_b200/csrc/edt_fh3.cuh) with a host context (one emulated thread per column, a tiny
"shared-memory" ring so that the local-memory backing is exercised too) and the result must be bit-identical
to the oracle's EDT for exactly representable anisotropies, within 1e-4 otherwise (the x pass multiplies
where the library adds repeatedly)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "host", "fh3_host.cpp")
HDR = os.path.join(ROOT, "kimimaro_b200", "csrc", "edt_fh3.cuh")
OUT = os.path.join(ROOT, "oracle", "_cache", "fh3_host.so")


@pytest.fixture(scope="module")
def fh3():
  os.makedirs(os.path.dirname(OUT), exist_ok=True)
  if (not os.path.exists(OUT)) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", SRC, "-o", OUT])
  return ctypes.CDLL(OUT)


def _host(lib, lab, an, bb, variant, hybrid=False):
  lab = np.asarray(lab)
  ndim = lab.ndim
  L = oracle._f(lab, np.uint32)
  sx, sy, sz = L.shape
  out = np.zeros(L.shape, np.float32, order="F")
  stats = (ctypes.c_long * 8)()
  an = tuple(float(a) for a in an) + (1.0,) * (3 - len(an))
  fn = lib.fh3_host_edt_hybrid if hybrid else lib.fh3_host_edt
  rc = fn(oracle._p(L), ctypes.c_int64(sx), ctypes.c_int64(sy), ctypes.c_int64(sz),
                        ctypes.c_float(an[0]), ctypes.c_float(an[1]), ctypes.c_float(an[2]), int(bool(bb)), ndim,
                        variant, oracle._p(out), stats)
  assert rc == 0
  return (out.reshape(lab.shape, order="F") if ndim == 2 else out), tuple(stats)


def _blocky(rng, shape, k, dense):
  lab = rng.integers(0, k + 1, size=shape).astype(np.uint32)
  rep = tuple(int(x) for x in rng.integers(1, 6, size=3))
  lab = np.repeat(np.repeat(np.repeat(lab[::rep[0], ::rep[1], ::rep[2]], rep[0], 0), rep[1], 1), rep[2], 2)
  lab = lab[:shape[0], :shape[1], :shape[2]]
  if dense:
    lab[lab == 0] = k + 1
  return np.asfortranarray(lab)


# variant = (ring entries C, rows between flushes R, load batch B), see fh3_host.cpp
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
def test_bit_identical_random(fh3, variant):
  rng = np.random.default_rng(100 + variant)
  for trial in range(14):
    shape = tuple(int(x) for x in rng.integers(1, 60, size=3))
    if trial % 5 == 0:
      shape = (int(rng.integers(1, 24)), int(rng.integers(100, 280)), int(rng.integers(1, 4)))
    lab = _blocky(rng, shape, int(rng.integers(1, 5)), dense=(trial % 3 == 0))
    for an in ((1, 1, 1), (16, 16, 40), (4, 4, 40)):
      for bb in (False, True):
        got, _ = _host(fh3, lab, an, bb, variant)
        assert np.array_equal(got, oracle.edt(lab, an, bb)), (shape, an, bb)


def test_known_shapes(fh3):
  plane = np.zeros((257, 257), np.uint32, order="F")
  plane[1:-1, 1:-1] = 1
  got, _ = _host(fh3, plane, (100, 100), True, 0)          # automated_test.py:104-114
  assert np.array_equal(got, oracle.edt(plane, (100, 100), True))
  assert np.unravel_index(np.argmax(got), got.shape) == (128, 128)
  ones = np.ones((40, 50, 60), np.uint32, order="F")
  for bb in (True, False):
    got, _ = _host(fh3, ones, (1, 1, 1), bb, 0)
    assert np.array_equal(got, oracle.edt(ones, (1, 1, 1), bb))
  assert np.isinf(got).all()
  zeros = np.zeros((17, 9, 4), np.uint32, order="F")
  assert not _host(fh3, zeros, (1, 2, 3), True, 1)[0].any()


def test_tubes_and_spill(fh3):
  from tests.synth import synthetic_tubes
  lab = synthetic_tubes((160, 128, 64), 25, seed=77)
  lab[40:120, 20:110, 8:56] = 999                            # a blob: deep stacks
  for variant in (0, 2):
    got, stats = _host(fh3, lab, (16, 16, 40), False, variant)
    assert np.array_equal(got, oracle.edt(lab, (16, 16, 40), False))
    assert stats[0] > 0                                      # the local-memory backing was exercised


def test_non_integer_anisotropy(fh3):
  from tests.synth import synthetic_tubes
  lab = synthetic_tubes((96, 80, 48), 12, seed=5)
  an = (3.3, 4.7, 10.1)
  got, _ = _host(fh3, lab, an, False, 0)
  ref = oracle.edt(lab, an, False)
  assert np.array_equal(np.isinf(got), np.isinf(ref)) and np.array_equal(got == 0, ref == 0)
  fin = np.isfinite(ref)
  np.testing.assert_allclose(got[fin], ref[fin], rtol=1e-4, atol=0)


# ---- hybrid pass (b2t_edt_ws): stencil everywhere + envelope on the flagged blocks, out of place ----
# variant = stencil windows (y, z) and prefetch, see fh3_host.cpp; every voxel of the result must have been
# written by one of the two kernels (the harness poisons both buffers with NaN)
@pytest.fixture(params=[(0, 1), (1, 1), (1, 4), (1, 5)], ids=["stencil_v1", "stencil_v2", "stencil_v2_qp4_pp", "stencil_v2_pp"])
def stencil(fh3, request):
  """The stencil bodies of edt_fh3.cuh -- stencil_column and stencil_column_v2 (leaner steady state) -- and the
  envelope's write-out with four entries of lookahead and the pop-ahead of its build (QP = 4, PP = 1 of column_range)."""
  fh3.fh3_host_set_stencil(request.param[0])
  fh3.fh3_host_set_query_prefetch(request.param[1])
  yield request.param
  fh3.fh3_host_set_stencil(0)
  fh3.fh3_host_set_query_prefetch(1)


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4])
def test_hybrid_bit_identical_random(fh3, variant, stencil):
  rng = np.random.default_rng(200 + variant)
  for trial in range(12):
    shape = tuple(int(x) for x in rng.integers(1, 60, size=3))
    if trial % 5 == 0:
      shape = (int(rng.integers(1, 24)), int(rng.integers(100, 280)), int(rng.integers(1, 4)))
    lab = _blocky(rng, shape, int(rng.integers(1, 5)), dense=(trial % 3 == 0))
    for an in ((1, 1, 1), (16, 16, 40), (4, 4, 40), (40, 32, 20)):
      for bb in (False, True):
        got, _ = _host(fh3, lab, an, bb, variant, hybrid=True)
        assert np.array_equal(got, oracle.edt(lab, an, bb)), (shape, an, bb)


def test_hybrid_known_shapes_and_blobs(fh3, stencil):
  plane = np.zeros((257, 257), np.uint32, order="F")
  plane[1:-1, 1:-1] = 1
  got, _ = _host(fh3, plane, (100, 100), True, 0, hybrid=True)
  assert np.array_equal(got, oracle.edt(plane, (100, 100), True))
  ones = np.ones((40, 50, 60), np.uint32, order="F")
  for bb in (True, False):
    got, _ = _host(fh3, ones, (1, 1, 1), bb, 0, hybrid=True)
    assert np.array_equal(got, oracle.edt(ones, (1, 1, 1), bb))
  from tests.synth import synthetic_tubes
  lab = synthetic_tubes((160, 128, 64), 25, seed=77)
  lab[40:120, 20:110, 8:56] = 999
  for variant in (0, 1):
    got, stats = _host(fh3, lab, (16, 16, 40), False, variant, hybrid=True)
    assert np.array_equal(got, oracle.edt(lab, (16, 16, 40), False))
    assert stats[3] > 0                                      # the envelope kernel had flagged blocks to redo


# ---- roles form (b2t_edt_config_roles): the previous pass predicts the stencil's misses; an envelope role over the
# prediction and a stencil role that skips it write the same buffer concurrently on the device, then the envelope over
# what the stencil flagged.  The harness runs the two roles in either order and with wrong predictions: the result must
# not depend on any of it (a prediction decides who computes a voxel, never what the value is).
def _roles(lib, lab, an, bb, variant, order, garbage=0):
  lab = np.asarray(lab)
  ndim = lab.ndim
  L = oracle._f(lab, np.uint32)
  sx, sy, sz = L.shape
  out = np.zeros(L.shape, np.float32, order="F")
  stats = (ctypes.c_long * 8)()
  an = tuple(float(a) for a in an) + (1.0,) * (3 - len(an))
  rc = lib.fh3_host_edt_roles(oracle._p(L), ctypes.c_int64(sx), ctypes.c_int64(sy), ctypes.c_int64(sz),
                              ctypes.c_float(an[0]), ctypes.c_float(an[1]), ctypes.c_float(an[2]), int(bool(bb)), ndim,
                              variant, order, garbage, oracle._p(out), stats)
  assert rc == 0
  return (out.reshape(lab.shape, order="F") if ndim == 2 else out), tuple(stats)


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4])
def test_roles_bit_identical_random(fh3, variant, stencil):
  rng = np.random.default_rng(300 + variant)
  for trial in range(10):
    shape = tuple(int(x) for x in rng.integers(1, 70, size=3))
    if trial % 5 == 0:
      shape = (int(rng.integers(1, 40)), int(rng.integers(100, 280)), int(rng.integers(1, 4)))
    lab = _blocky(rng, shape, int(rng.integers(1, 5)), dense=(trial % 3 == 0))
    for an in ((1, 1, 1), (16, 16, 40), (4, 4, 40), (40, 32, 20)):
      for bb in (False, True):
        ref = oracle.edt(lab, an, bb)
        for order in (0, 1):
          got, _ = _roles(fh3, lab, an, bb, variant, order, garbage=(trial + order) % 4)
          assert np.array_equal(got, ref), (shape, an, bb, order, trial % 4)


def test_roles_known_shapes_and_blobs(fh3, stencil):
  plane = np.zeros((257, 257), np.uint32, order="F")
  plane[1:-1, 1:-1] = 1
  for order in (0, 1):
    got, _ = _roles(fh3, plane, (100, 100), True, 0, order)
    assert np.array_equal(got, oracle.edt(plane, (100, 100), True))
  ones = np.ones((40, 50, 60), np.uint32, order="F")
  for bb in (True, False):
    got, _ = _roles(fh3, ones, (1, 1, 1), bb, 0, 1)
    assert np.array_equal(got, oracle.edt(ones, (1, 1, 1), bb))
  from tests.synth import synthetic_tubes
  lab = synthetic_tubes((160, 128, 64), 25, seed=77)
  lab[40:120, 20:110, 8:56] = 999
  ref = oracle.edt(lab, (16, 16, 40), False)
  for variant in (0, 1):
    for order in (0, 1):
      got, stats = _roles(fh3, lab, (16, 16, 40), False, variant, order)
      assert np.array_equal(got, ref)
      # the prediction found the blob (envelope warps had rows to do) and left the stencil little to flag
      assert stats[3] > 0 and stats[5] > 0
      assert stats[4] <= stats[5] // 4, stats
    for garbage in (1, 2, 3):
      got, _ = _roles(fh3, lab, (16, 16, 40), False, variant, 0, garbage)
      assert np.array_equal(got, ref)
