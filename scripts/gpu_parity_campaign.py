"""Randomised parity campaign on the GPU: random volumes (shapes with odd extents, several anisotropies including a
non-integer one that takes the envelope EDT, label dtypes, options) through kimimaro_b200.skeletonize in the default
order against the oracle in the same order, and in strict mode against the oracle's literal heap order (== the
reference's compiled extension).  One JSON line per case, a summary line at the end.
  python scripts/gpu_parity_campaign.py [cases=24] [seed=1]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import kimimaro_b200
from kimimaro_b200 import _lib
from kimimaro_b200.datasets import synthetic_tubes
from oracle import teasar

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
ANISO = [(16, 16, 40), (1, 1, 1), (4, 4, 40), (40, 32, 20), (8.5, 8.5, 33.0)]
DTYPES = [np.uint32, np.uint8, np.uint16, np.uint64, np.int32]


def same(a, b, rtol=1e-4):
  if sorted(a) != sorted(b):
    return False
  for k in b:
    if not (np.array_equal(a[k].vertices, b[k]["vertices"]) and np.array_equal(a[k].edges, b[k]["edges"])
            and np.allclose(a[k].radii, b[k]["radii"], rtol=rtol)):
      return False
  return True


ok_default = ok_strict = n_skel = 0
t_all = time.time()
for c in range(n_cases):
  shape = tuple(int(v) for v in rng.integers(40, 112, size=3))
  n_tubes = int(rng.integers(3, 14))
  an = ANISO[int(rng.integers(0, len(ANISO)))]
  dt = DTYPES[int(rng.integers(0, len(DTYPES)))]
  dense = c % 5 == 4          # every fifth case: a dense segmentation (Voronoi cells, no background), like real EM labels
  if dense:
    shape = tuple(int(v) for v in rng.integers(32, 72, size=3))
    pts = rng.random((n_tubes + 4, 3)) * np.array(shape)
    g = np.stack(np.meshgrid(*[np.arange(s_) for s_ in shape], indexing="ij"), axis=-1).astype(np.float32)
    d2 = ((g[..., None, :] - pts[None, None, None, :, :].astype(np.float32)) ** 2 * np.array(an, np.float32) ** 2).sum(-1)
    lab = np.asfortranarray((np.argmin(d2, axis=-1) + 1).astype(dt))
  else:
    lab = synthetic_tubes(shape, n_tubes, seed=int(rng.integers(1, 1 << 30)), anisotropy=an).astype(dt)
  kw = dict(anisotropy=an, dust_threshold=int(rng.choice([50, 200, 1000])), fix_borders=bool(rng.integers(0, 2)),
            fix_branching=bool(rng.integers(0, 4) > 0), fill_holes=bool(rng.integers(0, 5) == 0))
  if rng.integers(0, 3) == 0:
    kw["teasar_params"] = {"scale": float(rng.choice([1.0, 1.5, 4.0])), "const": float(rng.choice([0, 20, 300])),
                           "pdrf_scale": 100000, "pdrf_exponent": int(rng.choice([4, 8, 16]))}
  ref = teasar.skeletonize(lab, **kw)
  got = kimimaro_b200.skeletonize(lab, progress=False, **kw)
  d_ok = same(got, ref)
  mode, window = _lib.invalidation_mode()
  _lib.set_invalidation_mode("strict")
  try:
    got_s = kimimaro_b200.skeletonize(lab, progress=False, **kw)
  finally:
    _lib.set_invalidation_mode(mode, window)
  ref_s = teasar.skeletonize(lab, invalidation_mode="heap", **kw)
  s_ok = same(got_s, ref_s)
  ok_default += d_ok
  ok_strict += s_ok
  n_skel += len(ref)
  print(json.dumps({"case": c, "shape": shape, "tubes": n_tubes, "dense": bool(dense), "anisotropy": an, "dtype": np.dtype(dt).name,
                    "options": {k: v for k, v in kw.items() if k != "anisotropy"}, "skeletons": len(ref),
                    "default_equals_oracle": bool(d_ok), "strict_equals_reference_heap_order": bool(s_ok)}), flush=True)
print(json.dumps({"summary": True, "cases": n_cases, "seed": seed, "skeletons": n_skel, "default_equals_oracle": ok_default,
                  "strict_equals_reference_heap_order": ok_strict, "seconds": round(time.time() - t_all, 1)}), flush=True)
sys.exit(0 if ok_default == n_cases and ok_strict == n_cases else 1)
