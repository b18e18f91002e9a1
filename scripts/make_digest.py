"""tests/golden/synth512_oracle_digest_<mode>.json for one claim order of the oracle (CPU, all cores, a few minutes):
per skeleton the first 16 hex digits of sha256(vertices bytes + edges bytes).  Mode 'heap' is the reference's own order
(== its compiled extension voxel for voxel, tests/test_oracle_cpu.py::test_invalidation_vs_reference_ext): the digest the
engine's strict mode must hit on all skeletons and against which bench.py counts the default mode's agreement.

  python scripts/make_digest.py heap [512]"""
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

mode = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
vol = make_volume(n)
t = time.time()
sk = teasar.skeletonize(vol, anisotropy=ANISOTROPY, parallel=os.cpu_count(), invalidation_mode=mode)
print(mode, "oracle", round(time.time() - t, 1), "s", len(sk), "skeletons", flush=True)
dig = {}
for k, s in sk.items():
  h = hashlib.sha256()
  h.update(np.ascontiguousarray(s["vertices"]).tobytes())
  h.update(np.ascontiguousarray(s["edges"]).tobytes())
  dig[str(k)] = h.hexdigest()[:16]
name = mode.replace(":", "").replace(".", "p")
stem = "synth512" if n == 512 else f"synth{n}"
with open(os.path.join(ROOT, "tests", "golden", f"{stem}_oracle_digest_{name}.json"), "w") as f:
  json.dump({"volume": f"bench.make_volume({n})", "invalidation_mode": mode, "n_skeletons": len(sk),
             "n_vertices": int(sum(s["vertices"].shape[0] for s in sk.values())),
             "sha256_16_of_vertices_then_edges": dig}, f, indent=0)
