"""Full-size parity: the CUDA path vs the CPU oracle on the whole bench volume (synthetic-512).

  python scripts/full_parity.py oracle   # CPU only: runs the oracle (all cores), caches oracle/_cache/*.npz
  python scripts/full_parity.py gpu      # GPU box: runs kimimaro_b200 twice (determinism) and compares with the cache
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bench import make_volume, ANISOTROPY

n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
cache = os.path.join(ROOT, "oracle", "_cache", f"synth{n}_oracle.npz")
vol = make_volume(n)

if sys.argv[1] == "oracle":
  from oracle import teasar
  t = time.time()
  sk = teasar.skeletonize(vol, anisotropy=ANISOTROPY, parallel=os.cpu_count())
  print("oracle", time.time() - t, "s", len(sk), "skeletons")
  out = {"ids": np.array(sorted(sk), dtype=np.int64)}
  for k, s in sk.items():
    out[f"v{k}"] = s["vertices"]; out[f"e{k}"] = s["edges"]; out[f"r{k}"] = s["radii"]
  np.savez(cache, **out)
else:
  import kimimaro_b200
  a = kimimaro_b200.skeletonize(vol, anisotropy=ANISOTROPY, progress=False)
  b = kimimaro_b200.skeletonize(vol, anisotropy=ANISOTROPY, progress=False)
  same = sorted(a) == sorted(b) and all(np.array_equal(a[k].vertices, b[k].vertices) and np.array_equal(a[k].edges, b[k].edges) for k in a)
  print("deterministic across two runs:", same, len(a), "skeletons,", sum(s.vertices.shape[0] for s in a.values()), "vertices")
  if os.path.exists(cache):
    g = np.load(cache)
    ids = [int(i) for i in g["ids"]]
    print("ids equal:", sorted(a) == sorted(ids), len(ids))
    bad_v, bad_e, bad_r = [], [], []
    for k in ids:
      if k not in a:
        bad_v.append(k); continue
      if not np.array_equal(a[k].vertices, g[f"v{k}"]): bad_v.append(k)
      elif not np.array_equal(a[k].edges, g[f"e{k}"]): bad_e.append(k)
      elif not np.allclose(a[k].radii, g[f"r{k}"], rtol=1e-4): bad_r.append(k)
    print("labels with vertex mismatch:", len(bad_v), bad_v[:10], "| edge mismatch:", len(bad_e), "| radius mismatch:", len(bad_r))
    for k in bad_v[:3]:
      if k in a:
        print(" label", k, "cuda", a[k].vertices.shape, "oracle", g[f"v{k}"].shape)
  else:
    print("no oracle cache at", cache)
