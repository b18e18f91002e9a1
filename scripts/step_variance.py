"""Diagnostic: per-step time with GC and CUDA-allocator counters."""
import gc, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import make_volume, ANISOTROPY
from kimimaro_b200.intake import skeletonize
vol = make_volume(512)
d = torch.from_numpy(vol.reshape(-1, order="F").view(np.int32)).cuda()
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
if mode == "nogc":
  gc.disable()
for i in range(10):
  g0 = [s["collections"] for s in gc.get_stats()]
  m0 = torch.cuda.memory_stats()
  torch.cuda.synchronize(); t = time.perf_counter()
  sk = skeletonize(vol.shape, device_labels=d, anisotropy=ANISOTROPY, progress=False)
  torch.cuda.synchronize(); dt = time.perf_counter() - t
  g1 = [s["collections"] for s in gc.get_stats()]
  m1 = torch.cuda.memory_stats()
  print(i, round(dt * 1e3, 1), "gc", [b - a for a, b in zip(g0, g1)],
        "cudaMalloc", m1["num_device_alloc"] - m0["num_device_alloc"], "cudaFree", m1["num_device_free"] - m0["num_device_free"],
        "retries", m1["num_alloc_retries"] - m0["num_alloc_retries"], "reserved GB", round(m1["reserved_bytes.all.current"] / 2**30, 1), flush=True)
