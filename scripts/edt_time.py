"""Time b2t_edt alone (CUDA events) on a seeded synthetic volume; prints achieved algorithmic GB/s."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kimimaro_b200 import ops, _lib
from kimimaro_b200.datasets import synthetic_tubes

def main():
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
  reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
  _lib.require_device()
  t = time.time()
  if n == 512:
    import bench
    vol = bench.make_volume(512)
  else:
    vol = synthetic_tubes((n, n, n), 40, seed=0xB2002124)
  print("gen", time.time() - t, "fg frac", float((vol != 0).mean()), flush=True)
  d = ops.to_device_f(vol)
  out = torch.empty(vol.size, dtype=torch.float32, device="cuda")
  flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
  for _ in range(3):
    ops.edt(d, vol.shape, (16, 16, 40), False, out=out)
  torch.cuda.synchronize()
  times = []
  for _ in range(reps):
    flush.zero_()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); ops.edt(d, vol.shape, (16, 16, 40), False, out=out); e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
  ms = float(np.median(times))
  V = vol.size
  alg = (3 * vol.dtype.itemsize + 20) * V
  digest = int(torch.nan_to_num(out, posinf=0.0).view(torch.int32).to(torch.int64).sum().item())
  print(json.dumps({"digest": digest, "n": n, "ms_median": ms, "ms_min": min(times), "alg_GBps": alg / ms / 1e6, "voxels": V}))

main()
