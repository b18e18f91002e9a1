"""Per-phase times of one pass over the benchmark volume (synchronising laps), for A/B runs through environment variables:
  B2T_EDF_LABELS=0|1  B2T_EDF_TEAM_MIN=...  python scripts/phase_times.py [size=512] [reps=3]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import kimimaro_b200
from bench import make_volume, anisotropy_of, skeleton_digests, golden_digest

if os.environ.get("B2T_TRACE_LIMIT"):        # resident path-loop CTAs per SM (b2t_set_launch_limits)
  from kimimaro_b200 import _lib, engine  # noqa: F401
  _lib.lib().b2t_set_launch_limits(0, int(os.environ["B2T_TRACE_LIMIT"]))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
vol = make_volume(n)
an = anisotropy_of(n)
kimimaro_b200.skeletonize(vol, anisotropy=an, progress=False)
best = None
for _ in range(reps):
  tm = {}
  torch.cuda.synchronize()
  t = time.perf_counter()
  sk = kimimaro_b200.skeletonize(vol, anisotropy=an, progress=False, timings=tm)
  torch.cuda.synchronize()
  dt = time.perf_counter() - t
  if best is None or dt < best[0]:
    best = (dt, tm)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(reps):
  kimimaro_b200.skeletonize(vol, anisotropy=an, progress=False)
torch.cuda.synchronize()
free = (time.perf_counter() - t) / reps
rec = {"env": {k: v for k, v in os.environ.items() if k.startswith("B2T_")}, "pass_ms_with_laps": round(1e3 * best[0], 2),
       "pass_ms": round(1e3 * free, 2),
       "phases_ms": {k: round(1e3 * v, 2) for k, v in best[1].items() if isinstance(v, float)}}
st = best[1].get("kernel_stats")
if st:
  s0 = st[0]
  top = np.argsort(-s0["stats"][:, 3].astype(np.int64))[:8]
  rec["slowest_labels"] = [{"job": int(i), "us": int(s0["stats"][i, 3]), "npaths": int(s0["npaths"][i]),
                           "rounds": int(s0["stats"][i, 1]), "relax": int(s0["stats"][i, 0]),
                           "invalidated": int(s0["stats"][i, 2])} for i in top]
  us = np.sort(s0["stats"][:, 3].astype(np.int64))[::-1]
  rec["label_us"] = {"n": int(us.size), "sum": int(us.sum()), "max": int(us[0]) if us.size else 0,
                     "top16_sum": int(us[:16].sum()), "top148_sum": int(us[:148].sum()),
                     "q": [int(v) for v in np.quantile(us, [0.5, 0.9, 0.99])] if us.size else []}
gold = golden_digest("synth512_oracle_digest_window1.json") if n == 512 else None
if gold:
  dg = skeleton_digests(sk)
  rec["identical_to_oracle_same_mode"] = int(sum(dg.get(k) == h for k, h in gold.items()))
  rec["of"] = len(gold)
print(json.dumps(rec), flush=True)
