"""Cost and reference agreement of the three claim orders on the benchmark volume (GPU):
  python scripts/mode_cost.py [size=512]
per mode: path-loop time (synchronising laps), whole pass, skeletons identical to the reference's heap order
(tests/golden/synth512_oracle_digest_heap.json) and to the oracle in the same mode."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import kimimaro_b200
from kimimaro_b200 import _lib
from bench import make_volume, anisotropy_of, skeleton_digests, golden_digest

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
vol = make_volume(n)
an = anisotropy_of(n)
heap = golden_digest("synth512_oracle_digest_heap.json") if n == 512 else None
own = {"rounds": golden_digest("synth512_oracle_digest.json"), "window": golden_digest("synth512_oracle_digest_window1.json"),
       "strict": heap} if n == 512 else {}
out = []
kimimaro_b200.skeletonize(vol, anisotropy=an, progress=False)      # warm-up
for mode in ("window", "rounds", "strict"):
  _lib.set_invalidation_mode(mode)
  best = None
  for rep in range(2):
    tm = {}
    torch.cuda.synchronize()
    t = time.perf_counter()
    sk = kimimaro_b200.skeletonize(vol, anisotropy=an, progress=False, timings=tm)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    if best is None or dt < best[0]:
      best = (dt, tm)
  dg = skeleton_digests(sk)
  rec = {"mode": mode, "pass_ms": round(1e3 * best[0], 2), "path_loop_ms": round(1e3 * best[1].get("paths", 0), 2),
         "soma_ms": round(1e3 * best[1].get("soma", 0), 2), "skeletons": len(sk)}
  if heap:
    rec["identical_to_reference_heap_order"] = int(sum(dg.get(k) == h for k, h in heap.items()))
    if own.get(mode):
      rec["identical_to_oracle_same_mode"] = int(sum(dg.get(k) == h for k, h in own[mode].items()))
    rec["of"] = len(heap)
  st = best[1].get("kernel_stats")
  if st:
    s0 = st[0]["stats"]
    rec["slowest_label_us"] = int(s0[:, 3].max())
  print(json.dumps(rec), flush=True)
  out.append(rec)
_lib.set_invalidation_mode("window", 1.0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"mode_cost_{n}.json"), "w") as f:
  json.dump(out, f, indent=1)
