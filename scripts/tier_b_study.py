"""Tier B at BASELINE.json's full size, on the CPU: the oracle's claim orders for roll_invalidation_ball_inside_component
against its literal form (mode 'heap' == the reference's compiled extension voxel for voxel) on the whole benchmark
volume.  Writes tests/golden/synth512_oracle_digest_<mode>.json for every mode it runs (the digest format of
test_full_size_512_against_oracle_digest), so that a change of the engine's claim order finds its golden file waiting.

  python scripts/tier_b_study.py [n=512] [modes=rounds,window:1]"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bench import make_volume, ANISOTROPY
from oracle import teasar

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["rounds", "window:1"]
vol = make_volume(n)


def run(mode):
  t = time.time()
  sk = teasar.skeletonize(vol, anisotropy=ANISOTROPY, parallel=os.cpu_count(), invalidation_mode=mode)
  print(mode, "oracle", round(time.time() - t, 1), "s", len(sk), "skeletons", flush=True)
  return sk


def digest(sk):
  out = {}
  for k, s in sk.items():
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(s["vertices"]).tobytes())
    h.update(np.ascontiguousarray(s["edges"]).tobytes())
    out[str(k)] = h.hexdigest()[:16]
  return out


ref = run("heap")
ref_sets = {k: {tuple(x) for x in s["vertices"].tolist()} for k, s in ref.items()}
summary = {"volume": f"synthetic-{n}", "skeletons": len(ref), "vertices": int(sum(len(v) for v in ref_sets.values())), "modes": {}}
for mode in modes:
  sk = run(mode)
  same = sum(1 for k in ref if k in sk and {tuple(x) for x in sk[k]["vertices"].tolist()} == ref_sets[k])
  diff = sum(len({tuple(x) for x in sk[k]["vertices"].tolist()} ^ ref_sets[k]) for k in ref if k in sk)
  summary["modes"][mode] = {"identical_skeletons": same, "vertices_in_symmetric_difference": diff,
                            "same_ids": sorted(sk) == sorted(ref)}
  print(mode, summary["modes"][mode], flush=True)
  if n == 512:
    name = mode.replace(":", "").replace(".", "p")
    with open(os.path.join(ROOT, "tests", "golden", f"synth512_oracle_digest_{name}.json"), "w") as f:
      json.dump({"volume": "bench.make_volume(512)", "invalidation_mode": mode, "n_skeletons": len(sk),
                 "n_vertices": int(sum(s["vertices"].shape[0] for s in sk.values())),
                 "sha256_16_of_vertices_then_edges": digest(sk)}, f, indent=0)
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", f"r01_tier_b_study_{n}.json"), "w") as f:
  json.dump(summary, f, indent=1)
print(json.dumps(summary))
