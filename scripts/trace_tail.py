"""Per-label cost distribution of the path-loop kernel on the bench volume (diagnostic)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import make_volume, ANISOTROPY
from kimimaro_b200.intake import skeletonize
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
vol = make_volume(n)
d = torch.from_numpy(vol.reshape(-1, order="F").view(np.int32)).cuda()
for it in range(2):
  tm = {}
  sk = skeletonize(vol.shape, device_labels=d, anisotropy=ANISOTROPY, progress=False, timings=tm)
ks = tm["kernel_stats"][0]
st, npaths, seg = ks["stats"], ks["npaths"], ks["segids"]
us = st[:, 3].astype(np.int64)
order = np.argsort(-us)
print("labels", len(us), "sum_ms", us.sum() / 1e3, "max_ms", us.max() / 1e3, "paths total", int(npaths.sum()))
print("top 10: (us, npaths, rounds, relax, invalidated)")
for i in order[:10]:
  print(int(us[i]), int(npaths[i]), int(st[i, 1]), int(st[i, 0]), int(st[i, 2]))
print("percentiles us", np.percentile(us, [50, 90, 99]).tolist())
print({k: round(1e3 * v, 2) for k, v in tm.items() if isinstance(v, float)})
