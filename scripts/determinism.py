"""Diagnostic: run the CUDA path several times on the bench volume and report which labels differ."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench import make_volume, ANISOTROPY
import kimimaro_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
vol = make_volume(n)
cache = os.path.join(ROOT, "oracle", "_cache", f"synth{n}_oracle.npz")
g = np.load(cache) if os.path.exists(cache) else None
runs = []
for r in range(reps):
  tm = {}
  sk = kimimaro_b200.skeletonize(vol, anisotropy=ANISOTROPY, progress=False, timings=tm)
  runs.append(sk)
  st = tm["kernel_stats"]
  print("run", r, "skeletons", len(sk), "vertices", sum(s.vertices.shape[0] for s in sk.values()),
        "paths main", int(st[0]["npaths"].sum()), "private", [int(x["npaths"].sum()) for x in st[1:]], flush=True)
  if g is not None:
    bad = [k for k in sk if not (np.array_equal(sk[k].vertices, g[f"v{k}"]) and np.array_equal(sk[k].edges, g[f"e{k}"]))]
    print("   vs oracle: mismatching labels", len(bad), bad[:8])
    for k in bad[:4]:
      a, b = sk[k].vertices, g[f"v{k}"]
      sa = set(map(tuple, a.tolist())); sb = set(map(tuple, b.tolist()))
      print("    label", k, "cuda", a.shape[0], "oracle", b.shape[0], "only cuda", len(sa - sb), "only oracle", len(sb - sa),
            "voxels in label", int((vol == k).sum()))
    rbad = []
    for k in sk:
      if k in bad: continue
      rel = np.abs(sk[k].radii - g[f"r{k}"]) / np.maximum(g[f"r{k}"], 1e-9)
      if rel.max() > 1e-4: rbad.append((k, float(rel.max()), int(np.argmax(rel)), float(sk[k].radii[np.argmax(rel)]), float(g[f"r{k}"][np.argmax(rel)])))
    print("   radius mismatches:", rbad[:5])
for r in range(1, reps):
  diff = [k for k in runs[0] if k not in runs[r] or not np.array_equal(runs[0][k].vertices, runs[r][k].vertices)]
  print("run 0 vs", r, "differ on", len(diff), diff[:8])
