"""A/B of the K1 forms WITHOUT torch (python + numpy + ctypes on libcudart / libb2t.so), for a GPU box with a minute to
spare: every form is checked bit for bit against the shipped default on the benchmark volume and on a few small shapes,
then timed with CUDA events (L2 flushed between repetitions).  One JSON object per line, also appended to
gpurun_out/k1_shot.jsonl as it goes so that a cut-off run keeps what it measured.

  python scripts/k1_shot.py [reps]
"""
import ctypes
import glob
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
T0 = time.time()
OUT = os.path.join(ROOT, "gpurun_out", "k1_shot.jsonl")
os.makedirs(os.path.dirname(OUT), exist_ok=True)


def say(**kw):
  kw["t"] = round(time.time() - T0, 2)
  line = json.dumps(kw)
  print(line, flush=True)
  with open(OUT, "a") as f:
    f.write(line + "\n")


c_vp, c_i64, c_int, c_f32, c_sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
rt = None
for cand in ("libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
  try:
    rt = ctypes.CDLL(cand, mode=ctypes.RTLD_GLOBAL)
    break
  except OSError:
    pass
b2t = ctypes.CDLL(os.environ.get("B2T_LIB") or os.path.join(ROOT, "kimimaro_b200", "libb2t.so"))
if rt is None:   # libb2t.so pulled its own libcudart in
  rt = ctypes.CDLL(None)
b2t.b2t_last_error.restype = ctypes.c_char_p
b2t.b2t_edt_workspace_bytes.restype = c_sz
b2t.b2t_edt_workspace_bytes.argtypes = [c_i64, c_i64, c_i64]
b2t.b2t_edt_config_roles.argtypes = [c_int, c_int, c_f32]
b2t.b2t_edt_ws.argtypes = [c_vp, c_int, c_i64, c_i64, c_i64, c_f32, c_f32, c_f32, c_int, c_int, c_vp, c_vp, c_sz, c_vp]
rt.cudaMalloc.argtypes = [ctypes.POINTER(c_vp), c_sz]
rt.cudaMemcpy.argtypes = [c_vp, c_vp, c_sz, c_int]
rt.cudaMemset.argtypes = [c_vp, c_int, c_sz]
rt.cudaMemsetAsync.argtypes = [c_vp, c_int, c_sz, c_vp]
rt.cudaEventCreate.argtypes = [ctypes.POINTER(c_vp)]
rt.cudaEventRecord.argtypes = [c_vp, c_vp]
rt.cudaEventSynchronize.argtypes = [c_vp]
rt.cudaEventElapsedTime.argtypes = [ctypes.POINTER(c_f32), c_vp, c_vp]
rt.cudaFree.argtypes = [c_vp]
rt.cudaGetErrorString.restype = ctypes.c_char_p


def cu(rc, what):
  if rc != 0:
    raise RuntimeError(f"{what}: cuda error {rc} {rt.cudaGetErrorString(rc)}")


def ok(rc, what):
  if rc != 0:
    raise RuntimeError(f"{what}: status {rc}: {b2t.b2t_last_error()}")


def dmalloc(n):
  p = c_vp()
  cu(rt.cudaMalloc(ctypes.byref(p), n), "cudaMalloc")
  return p


def h2d(dst, arr):
  cu(rt.cudaMemcpy(dst, arr.ctypes.data_as(c_vp), arr.nbytes, 1), "h2d")


def d2h(arr, src):
  cu(rt.cudaMemcpy(arr.ctypes.data_as(c_vp), src, arr.nbytes, 2), "d2h")


# (name, hybrid on, roles on, stencil_column_v2 on, prediction scale, (query prefetch, pop-ahead) of the envelope kernel)
# the first form is the reference the others are compared with: the shipped default
FORMS = [
  ("hybrid + stencil v2 (shipped default)", 1, 0, 1, 1.0, (1, 0)),
  ("hybrid + stencil v2 + envelope query prefetch 4", 1, 0, 1, 1.0, (4, 0)),
  ("hybrid + stencil v2 + envelope pop-ahead", 1, 0, 1, 1.0, (1, 1)),
  ("hybrid + stencil v2 + envelope query prefetch 4 + pop-ahead", 1, 0, 1, 1.0, (4, 1)),
  ("hybrid + stencil v1", 1, 0, 0, 1.0, (1, 0)),
  ("roles + stencil v2", 1, 1, 1, 1.0, (1, 0)),
  ("roles + stencil v2 + envelope query prefetch 4 + pop-ahead", 1, 1, 1, 1.0, (4, 1)),
  ("envelope only (b2t_edt path)", 0, 0, 1, 1.0, (1, 0)),
]


def edt(d_lab, shape, an, bb, ndim, d_out, d_ws, ws_bytes):
  sx, sy, sz = shape
  ok(b2t.b2t_edt_ws(d_lab, 4, sx, sy, sz, an[0], an[1], an[2], int(bb), ndim, d_out, d_ws, ws_bytes, None), "b2t_edt_ws")


def main():
  reps = int(sys.argv[1]) if len(sys.argv) > 1 else 7
  ok(b2t.b2t_device_check(), "device_check")
  say(event="start")
  packed = sorted(glob.glob(os.path.join(ROOT, "oracle", "_cache", "synth_512_*.npz")))
  if packed:
    vol = np.asfortranarray(np.load(packed[0])["v"].astype(np.uint32))
  else:
    import bench
    vol = bench.make_volume(512)
  say(event="volume", shape=list(vol.shape))
  an = (16.0, 16.0, 40.0)
  V = vol.size
  alg = 32 * V
  d_lab = dmalloc(V * 4)
  h2d(d_lab, vol.reshape(-1, order="F"))
  d_ref, d_out = dmalloc(V * 4), dmalloc(V * 4)
  ws_bytes = int(b2t.b2t_edt_workspace_bytes(*vol.shape))
  d_ws = dmalloc(ws_bytes)
  flush_bytes = 256 << 20
  d_flush = dmalloc(flush_bytes)
  e0, e1 = c_vp(), c_vp()
  cu(rt.cudaEventCreate(ctypes.byref(e0)), "event")
  cu(rt.cudaEventCreate(ctypes.byref(e1)), "event")
  h_ref = np.empty(V, np.float32)
  h_out = np.empty(V, np.float32)

  def timed(fn):
    ts = []
    for _ in range(reps):
      cu(rt.cudaMemsetAsync(d_flush, 0, flush_bytes, None), "flush")
      cu(rt.cudaEventRecord(e0, None), "rec")
      fn()
      cu(rt.cudaEventRecord(e1, None), "rec")
      cu(rt.cudaEventSynchronize(e1), "sync")
      ms = c_f32()
      cu(rt.cudaEventElapsedTime(ctypes.byref(ms), e0, e1), "elapsed")
      ts.append(ms.value)
    return ts

  # small shapes (2-D planes, borders, other anisotropies): every form against the envelope-only path
  from kimimaro_b200.datasets import synthetic_tubes
  small = []
  rng = np.random.default_rng(3)
  for shape, an2, bb in [((256, 192, 96), (16, 16, 40), False), ((256, 192, 96), (4, 4, 40), True),
                         ((128, 300, 40), (1, 1, 1), False), ((64, 64, 36), (40, 32, 20), True),
                         ((260, 257), (100, 100), True), ((512, 512), (16, 16), True)]:
    if len(shape) == 3:
      lab = synthetic_tubes(shape, 40, seed=int(rng.integers(1 << 30)))
      lab[shape[0] // 4: shape[0] // 2, shape[1] // 4: shape[1] // 2, shape[2] // 4: shape[2] // 2] = 7777
    else:
      lab = np.zeros(shape, np.uint32, order="F"); lab[1:-1, 1:-1] = 1; lab[100:140, 50:200] = 2
    small.append((np.asfortranarray(lab.astype(np.uint32)), an2, bb))

  def small_cases():
    res = []
    for lab, an2, bb in small:
      shp = tuple(lab.shape) + (1,) * (3 - lab.ndim)
      a3 = tuple(float(a) for a in an2) + (1.0,) * (3 - len(an2))
      n = lab.size
      h2d(d_flush, lab.reshape(-1, order="F"))                 # the flush buffer doubles as the small label volume
      wsb = int(b2t.b2t_edt_workspace_bytes(*shp))
      cu(rt.cudaMemset(d_out, 0xff, n * 4), "poison")
      cu(rt.cudaMemset(d_ws, 0xff, n * 4), "poison")
      edt(d_flush, shp, a3, bb, lab.ndim, d_out, d_ws, wsb)
      got = np.empty(n, np.float32)
      d2h(got, d_out)
      res.append(got)
    return res

  first = True
  small_ref = None
  for name, hybrid, roles, sv2, pscale, qp in FORMS:
    try:
      ok(b2t.b2t_edt_config_hybrid(hybrid, 0, 0, 4, 11, 8), "config_hybrid")
      ok(b2t.b2t_edt_config_roles(roles, sv2, c_f32(pscale)), "config_roles")
      ok(b2t.b2t_edt_config_envelope(qp[0], qp[1]), "config_envelope")
      dst = d_ref if first else d_out
      cu(rt.cudaMemset(dst, 0xff, V * 4), "poison")
      cu(rt.cudaMemset(d_ws, 0xff, V * 4), "poison")           # a voxel nobody writes stays a NaN
      edt(d_lab, vol.shape, an, False, 3, dst, d_ws, ws_bytes)
      cu(rt.cudaDeviceSynchronize(), "sync")
      rec = {"form": name}
      if first:
        d2h(h_ref, d_ref)
        rec["finite_or_inf_everywhere"] = bool(not np.isnan(h_ref).any())
      else:
        d2h(h_out, d_out)
        same = bool(np.array_equal(h_out, h_ref))
        rec["identical_to_default"] = same
        if not same:
          rec["mismatches"] = int((h_out != h_ref).sum())
      ts = timed(lambda: edt(d_lab, vol.shape, an, False, 3, dst, d_ws, ws_bytes))
      rec.update(ms_median=float(np.median(ts)), ms_min=float(min(ts)), ms_all=[round(t, 4) for t in ts],
                 alg_GBps=alg / float(np.median(ts)) / 1e6)
      sc = small_cases()
      if small_ref is None:
        small_ref = sc
        rec["small_cases_nan_free"] = [bool(not np.isnan(a).any()) for a in sc]
      else:
        rec["small_cases_identical"] = [bool(np.array_equal(a, b)) for a, b in zip(sc, small_ref)]
      h2d(d_flush, np.zeros(1, np.uint32))
    except Exception as e:   # a form that fails must not hide the others
      rec = {"form": name, "error": str(e)}
    first = False
    say(**rec)
  say(event="done")


main()
