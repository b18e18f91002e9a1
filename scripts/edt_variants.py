"""A/B of the K1 column-pass kernels on one GPU: every compiled variant is checked for bit-identity against
the v2 kernel's output (and the default against the CPU oracle) and timed with CUDA events, L2 flushed
between repetitions.  Prints one JSON object per variant and writes them to gpurun_out/edt_variants.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from kimimaro_b200 import ops, _lib

VARIANTS = [(16, 6, 32, 4)]


def timeit(fn, flush, reps):
  ts = []
  for _ in range(reps):
    flush.zero_()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
  return float(np.median(ts)), float(min(ts))


def main():
  reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
  check_oracle = "--oracle" in sys.argv
  _lib.require_device()
  lib = _lib.lib()
  vol = bench.make_volume(512)
  an = bench.ANISOTROPY
  V = vol.size
  alg = (3 * 4 + 20) * V
  d = ops.to_device_f(vol)
  flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
  res = []

  def run(name, out):
    ops.edt(d, vol.shape, an, False, out=out, workspace=False)
    torch.cuda.synchronize()

  ref = torch.empty(V, dtype=torch.float32, device="cuda")
  _lib.check(lib.b2t_edt_config(2, 0, 0, 0, 0))
  run("v2", ref)
  med, mn = timeit(lambda: ops.edt(d, vol.shape, an, False, out=ref, workspace=False), flush, reps)
  res.append({"variant": "v2 (local-memory F-H)", "ms_median": med, "ms_min": mn, "alg_GBps": alg / med / 1e6})
  print(json.dumps(res[-1]), flush=True)
  if check_oracle:
    from oracle import edt as orc_edt
    t = time.time()
    o = orc_edt(vol, an, False).reshape(-1, order="F")
    same = bool(np.array_equal(ref.cpu().numpy(), o))
    print(json.dumps({"v2_equals_oracle": same, "oracle_s": time.time() - t}), flush=True)
  # dense variant of the volume (every voxel labelled, like a real segmentation): background -> blocks of 24^3
  idx = np.indices((512 // 32, 512 // 32, 512 // 32)).reshape(3, -1)
  blk = (3000 + idx[0] + 16 * (idx[1] + 16 * idx[2])).reshape(16, 16, 16).astype(np.uint32)
  dense = np.asfortranarray(np.where(vol != 0, vol, np.kron(blk, np.ones((32, 32, 32), np.uint32))))
  dd = ops.to_device_f(dense)
  ref_d = torch.empty(V, dtype=torch.float32, device="cuda")
  ops.edt(dd, vol.shape, an, False, out=ref_d, workspace=False)
  medd, mnd = timeit(lambda: ops.edt(dd, vol.shape, an, False, out=ref_d, workspace=False), flush, max(3, reps // 2))
  res.append({"variant": "v2 dense", "ms_median": medd, "ms_min": mnd})
  print(json.dumps(res[-1]), flush=True)

  out = torch.empty(V, dtype=torch.float32, device="cuda")
  for (c, mb, r, b) in VARIANTS:
    _lib.check(lib.b2t_edt_config(3, c, mb, r, b))
    try:
      out.fill_(-1.0)
      run("v3", out)
      same = bool(torch.equal(out, ref))
      nbad = int((out != ref).sum().item()) if not same else 0
      med, mn = timeit(lambda: ops.edt(d, vol.shape, an, False, out=out, workspace=False), flush, reps)
      out.fill_(-1.0)
      ops.edt(dd, vol.shape, an, False, out=out, workspace=False)
      torch.cuda.synchronize()
      same_d = bool(torch.equal(out, ref_d))
      medd, mnd = timeit(lambda: ops.edt(dd, vol.shape, an, False, out=out, workspace=False), flush, max(3, reps // 2))
      rec = {"variant": f"v3 C={c} minb={mb} R={r} B={b}", "identical_to_v2": same, "mismatches": nbad,
             "ms_median": med, "ms_min": mn, "alg_GBps": alg / med / 1e6, "dense_identical": same_d,
             "dense_ms_median": medd}
    except Exception as e:  # a variant that fails must not hide the others
      rec = {"variant": f"v3 C={c} minb={mb} R={r} B={b}", "error": str(e)}
    res.append(rec)
    print(json.dumps(rec), flush=True)
  # hybrid (b2t_edt_ws): stencil windows (y, z), prefetch, min blocks per SM x envelope variant for the flagged blocks
  if "--hybrid" in sys.argv:
    HY = [((10, 4, 4, 11, 8), (16, 6, 32, 4)), ((8, 4, 4, 11, 8), (16, 6, 32, 4)), ((6, 4, 4, 11, 8), (16, 6, 32, 4)),
          ((10, 4, 6, 11, 8), (16, 6, 32, 4)), ((10, 4, 3, 11, 8), (16, 6, 32, 4)), ((10, 4, 4, 7, 8), (16, 6, 32, 4)),
          ((12, 4, 4, 7, 8), (16, 6, 32, 4)), ((10, 6, 4, 11, 8), (16, 6, 32, 4)), ((10, 4, 4, 11, 8), (32, 4, 32, 8)),
          ((10, 4, 4, 11, 8), (16, 8, 32, 4))]
    for (wy_, wz_, wr_, pf, hmb), (c, mb, r, b) in HY:
      _lib.check(lib.b2t_edt_config(3, c, mb, r, b))
      _lib.check(lib.b2t_edt_config_hybrid(1, wy_, wz_, wr_, pf, hmb))
      name = f"hybrid W=({wy_},{wz_}) wr={wr_} pf={pf} minb={hmb} + env C={c} minb={mb} R={r} B={b}"
      try:
        out.fill_(-1.0)
        ops.edt(d, vol.shape, an, False, out=out, workspace=True)
        torch.cuda.synchronize()
        same = bool(torch.equal(out, ref))
        nbad = int((out != ref).sum().item()) if not same else 0
        med, mn = timeit(lambda: ops.edt(d, vol.shape, an, False, out=out, workspace=True), flush, reps)
        out.fill_(-1.0)
        ops.edt(dd, vol.shape, an, False, out=out, workspace=True)
        torch.cuda.synchronize()
        same_d = bool(torch.equal(out, ref_d))
        medd, mnd = timeit(lambda: ops.edt(dd, vol.shape, an, False, out=out, workspace=True), flush, max(3, reps // 2))
        rec = {"variant": name, "identical_to_v2": same, "mismatches": nbad, "ms_median": med, "ms_min": mn,
               "alg_GBps": alg / med / 1e6, "dense_identical": same_d, "dense_ms_median": medd}
      except Exception as e:
        rec = {"variant": name, "error": str(e)}
      res.append(rec)
      print(json.dumps(rec), flush=True)
    # other shapes / borders / 2-D through the hybrid, against the in-place v2 kernels
    from kimimaro_b200.datasets import synthetic_tubes
    _lib.check(lib.b2t_edt_config_hybrid(1, 10, 4, 4, 11, 8))
    _lib.check(lib.b2t_edt_config(3, 16, 6, 32, 4))
    rng = np.random.default_rng(3)
    small = []
    for shape, an2, bb in [((256, 192, 96), (16, 16, 40), False), ((256, 192, 96), (4, 4, 40), True),
                           ((128, 300, 40), (1, 1, 1), False), ((64, 64, 33), (40, 32, 20), True),
                           ((260, 257), (100, 100), True), ((512, 512), (16, 16), True)]:
      if len(shape) == 3:
        lab = synthetic_tubes(shape, 40, seed=int(rng.integers(1 << 30)))
        lab[shape[0] // 4: shape[0] // 2, shape[1] // 4: shape[1] // 2, shape[2] // 4: shape[2] // 2] = 7777
      else:
        lab = np.zeros(shape, np.uint32, order="F"); lab[1:-1, 1:-1] = 1; lab[100:140, 50:200] = 2
      dl = ops.to_device_f(lab)
      _lib.check(lib.b2t_edt_config(2, 0, 0, 0, 0))
      r2 = ops.edt(dl, lab.shape, an2, bb, workspace=False).clone()
      _lib.check(lib.b2t_edt_config(3, 16, 6, 32, 4))
      r3 = ops.edt(dl, lab.shape, an2, bb, workspace=True)
      torch.cuda.synchronize()
      small.append({"shape": list(shape), "an": list(an2), "bb": bb, "identical": bool(torch.equal(r2, r3))})
    print(json.dumps({"hybrid_small_cases": small}), flush=True)
    res.append({"hybrid_small_cases": small})
  os.makedirs("gpurun_out", exist_ok=True)
  with open("gpurun_out/edt_variants.json", "w") as f:
    json.dump(res, f, indent=1)


main()
