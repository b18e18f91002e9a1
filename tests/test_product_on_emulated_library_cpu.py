"""The PRODUCT on the CPU suite: kimimaro_b200.skeletonize -- its whole Python host path (intake.py, engine.py,
border.py, skeleton.py) -- runs here on CPU tensors against oracle/_cache/libb2t_emu.so, the library's own kernels
compiled for the CPU against the SIMT emulation (tests/host/*_emu.cpp; K1 through the host context of its column-pass
templates), and must give the oracle's skeletons bit for bit, like the GPU tests ask of the real library.

This is test infrastructure: the product is not touched -- the test swaps the library path and stubs the handful of
torch.cuda calls the host code makes (tensor.cuda(), current_stream, is_available); without those stubs the product
still refuses to run without an sm_100 device (tests/test_oracle_cpu.py::test_no_cpu_fallback)."""
import contextlib
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HOST = os.path.join(HERE, "host")
UNITS = ["trace_emu", "preamble_emu", "field_emu", "b2t_emu_main"]
OUT = os.path.join(ROOT, "oracle", "_cache", "libb2t_emu.so")


def _build():
  deps = [os.path.join(HOST, u + ".cpp") for u in UNITS] + [os.path.join(HOST, "fh3_host.cpp")]
  deps += [os.path.join(HOST, "emu_include", f) for f in os.listdir(os.path.join(HOST, "emu_include"))]
  deps += [os.path.join(ROOT, "kimimaro_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "kimimaro_b200", "csrc"))]
  if os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in deps):
    return OUT
  obj_dir = os.path.join(ROOT, "oracle", "_cache", "emuobj")
  os.makedirs(obj_dir, exist_ok=True)
  flags = ["-O1", "-ffp-contract=off", "-std=c++17", "-fPIC", "-pthread", "-Wno-attributes", "-DB2T_EMU_COMBINED",
           "-I" + os.path.join(HOST, "emu_include"), "-I" + HOST]
  objs = []
  for u in UNITS:
    o = os.path.join(obj_dir, u + ".o")
    subprocess.check_call(["g++"] + flags + ["-c", os.path.join(HOST, u + ".cpp"), "-o", o])
    objs.append(o)
  subprocess.check_call(["g++", "-shared", "-pthread"] + objs + ["-o", OUT])
  return OUT


class _Stream:
  cuda_stream = 0


@pytest.fixture()
def product(monkeypatch):
  import torch
  from kimimaro_b200 import _lib
  import kimimaro_b200.intake  # noqa: F401  (registers its entry points with _lib.declare)
  import kimimaro_b200.engine  # noqa: F401
  lib_path = _build()
  monkeypatch.setattr(_lib, "LIB_PATH", lib_path)
  monkeypatch.setattr(_lib, "_lib", None)
  monkeypatch.setattr(_lib, "_checked_devices", set())
  monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
  monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
  monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _Stream())
  monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
  monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
  monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
  monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True), raising=False)   # ops.edt asserts it
  import kimimaro_b200
  yield kimimaro_b200
  monkeypatch.undo()


def _compare(res, ref):
  assert sorted(res.keys()) == sorted(ref.keys())
  for k in ref:
    a, b = res[k], ref[k]
    assert np.array_equal(a.vertices, b["vertices"]), k
    assert np.array_equal(a.edges, b["edges"]), k
    np.testing.assert_allclose(a.radii, b["radii"], rtol=1e-4)


def test_sphere_and_tubes(product):
  from oracle import teasar
  from tests.synth import sphere, synthetic_tubes
  for labels, kw in ((sphere(40, 14), {"dust_threshold": 100}),
                     (synthetic_tubes((56, 48, 32), 5, seed=3), {"anisotropy": (16, 16, 40), "dust_threshold": 100})):
    res = product.skeletonize(labels, progress=False, **kw)
    ref = teasar.skeletonize(labels, **kw)
    assert len(ref) >= 1
    _compare(res, ref)


def _small_cell_with_nucleus():
  v = np.zeros((48, 40, 32), np.uint32, order="F")
  v[4:44, 10:30, 6:26] = 5           # a cell body ...
  v[15:25, 15:25, 12:20] = 9         # ... whose nucleus is a label of its own
  v[30:35, 17:22, 14:18] = 0         # ... and a vacuole (an enclosed void)
  v[4:44, 33:38, 10:20] = 7          # a second, solid process
  return v


def test_fill_holes_and_fix_avocados(product):
  """The options of SURVEY 8f N4 through the product's real pipeline (tests/test_zz_options_gpu.py asks the same of the
  GPU): nucleus swallowed, vacuole filled, skeletons equal to the oracle's."""
  from oracle import teasar
  labels = _small_cell_with_nucleus()
  base = dict(anisotropy=(16, 16, 40), dust_threshold=50)
  ref0 = teasar.skeletonize(labels, **base)
  assert sorted(ref0) == [5, 7, 9]
  _compare(product.skeletonize(labels, progress=False, **base), ref0)
  kw = dict(base, fill_holes=True)
  ref = teasar.skeletonize(labels, **kw)
  assert sorted(ref) == [5, 7]
  _compare(product.skeletonize(labels, progress=False, **kw), ref)
  params = dict(product.DEFAULT_TEASAR_PARAMS)
  params["soma_detection_threshold"] = 150        # candidates: DBF above 150 / 2.5 nm (intake.py:619)
  kw = dict(base, fix_avocados=True, teasar_params=params)
  ref = teasar.skeletonize(labels, **kw)
  assert sorted(ref) == [5, 7]
  _compare(product.skeletonize(labels, progress=False, **kw), ref)
  kw["fill_holes"] = True
  _compare(product.skeletonize(labels, progress=False, **kw), teasar.skeletonize(labels, **kw))
