"""Generates tests/golden/golden_v1.npz: skeletons of two seeded cases computed by the CPU oracle
(oracle/teasar.py with the compiled C restatement).  The reference package itself cannot be imported
in this image (its PyPI dependencies are absent), so these vectors freeze the ORACLE, whose pieces are
pinned separately against the reference's in-tree extension and known-answer tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import teasar  # noqa: E402
from kimimaro_b200.datasets import sphere, synthetic_tubes  # noqa: E402

cases = {"sphere": (sphere(64, 24), {}),
         "tubes": (synthetic_tubes((96, 96, 64), 12, seed=1), {"anisotropy": (16, 16, 40), "dust_threshold": 100})}
# one file per claim order the engine can run (oracle/teasar.py: DEFAULT_INVALIDATION_MODE names the current one);
# tests/golden_name() picks the file
# "heap" = the reference's own order (the engine's strict mode)
for mode, fname in (("rounds", "golden_v1.npz"), ("window:1", "golden_v1_window1.npz"), ("heap", "golden_v1_heap.npz")):
  out = {}
  for name, (lab, kw) in cases.items():
    sk = teasar.skeletonize(lab, invalidation_mode=mode, **kw)
    out[f"{name}_ids"] = np.array(sorted(sk), dtype=np.int64)
    for i, s in sk.items():
      out[f"{name}_{i}_v"] = s["vertices"]
      out[f"{name}_{i}_e"] = s["edges"]
      out[f"{name}_{i}_r"] = s["radii"]
  np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), fname), **out)
  print(mode, "wrote", len(out), "arrays to", fname)
