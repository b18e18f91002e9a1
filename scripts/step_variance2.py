"""Diagnostic: which phase is slow on outlier steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import make_volume, ANISOTROPY
from kimimaro_b200.intake import skeletonize
vol = make_volume(512)
d = torch.from_numpy(vol.reshape(-1, order="F").view(np.int32)).cuda()
rows = []
for i in range(16):
  tm = {}
  torch.cuda.synchronize(); t = time.perf_counter()
  skeletonize(vol.shape, device_labels=d, anisotropy=ANISOTROPY, progress=False, timings=tm)
  torch.cuda.synchronize(); dt = time.perf_counter() - t
  ks = tm["kernel_stats"][0]
  rows.append((dt * 1e3, {k: round(v * 1e3, 1) for k, v in tm.items() if isinstance(v, float)}, int(ks["stats"][:, 3].max()), int(ks["stats"][:, 0].sum()), int(ks["stats"][:, 1].sum())))
rows = rows[1:]
med = np.median([r[0] for r in rows])
keys = [k for k in rows[0][1] if not k.startswith("soma_")]
print("median step", round(med, 1))
print("median phases", {k: float(np.median([r[1][k] for r in rows])) for k in keys})
for r in rows:
  flag = "SLOW" if r[0] > med * 1.1 else "    "
  print(flag, round(r[0], 1), "max label us", r[2], "relax", r[3], "rounds", r[4], {k: r[1][k] for k in keys if r[1][k] > 1.3 * np.median([q[1][k] for q in rows]) + 1})
