"""CUDA-event time of engine.connected_components on the benchmark volume (and on a dense relabelling of it):
  B2T_LIB=... python scripts/ccl_time.py [size=512] [reps=5]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from kimimaro_b200 import engine
from bench import make_volume

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
vol = make_volume(n)
rec = {"env": {k: v for k, v in os.environ.items() if k.startswith("B2T_")}}
for name, v in (("synthetic", vol), ("dense", vol + np.uint32(1))):    # dense: the background becomes one more label
  d = torch.from_numpy(v.reshape(-1, order="F").view(np.int32)).cuda()
  cc, n_cc = engine.connected_components(d, vol.shape)
  ts = []
  for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    cc, n_cc = engine.connected_components(d, vol.shape)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  rec[name] = {"ms": round(float(np.median(ts)), 3), "n_cc": n_cc, "checksum": int(cc.to(torch.int64).sum().item())}
print(json.dumps(rec), flush=True)
