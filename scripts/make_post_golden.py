#!/usr/bin/env python
"""
TEST INFRASTRUCTURE -- generates tests/golden/post_golden_v2.npz by RUNNING the reference's own
/root/reference/kimimaro/post.py (unmodified, loaded from where it lies) on seeded random merged skeletons.

The reference module imports four things this image does not have; they are supplied as follows and nothing else is
substituted:
  kimimaro.skeletontricks   the reference's own extension, compiled into oracle/_ref by oracle/build_ref.py
                            (find_cycle, create_distance_graph are the real C++)
  osteoid.Skeleton          kimimaro_b200.skeleton.Skeleton (osteoid is not installed; the class restates the subset of
                            its behaviour the reference touches: components, consolidate, simple_merge, cable_length)
  fastremap.unique          numpy.unique (same contract: sorted unique values, optional counts)
  networkx                  a 40-line stand-in below: Graph.add_edges_from / remove_edges_from / edges and shortest_path
                            (breadth first; on the trees remove_ticks works on the path is unique)
scipy's cKDTree is what the reference itself falls back to without pykdtree (post.py:39-42).

Run here (needs /root/reference):  python scripts/make_post_golden.py
tests/test_post_cpu.py::test_against_reference_run_goldens replays the inputs through kimimaro_b200.post.
"""
import importlib.util
import os
import sys
import types
from collections import deque

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("KIMIMARO_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "post_golden_v2.npz")


def load_reference_post():
  from oracle import build_ref
  build_ref.build(verbose=False)
  ext = build_ref.load()
  from kimimaro_b200.skeleton import Skeleton

  class Graph:
    def __init__(self):
      self.adj = {}

    def add_edges_from(self, edges):
      for a, b in edges:
        a, b = int(a), int(b)
        self.adj.setdefault(a, {})[b] = None
        self.adj.setdefault(b, {})[a] = None

    def remove_edges_from(self, edges):
      for a, b in edges:
        self.adj.get(a, {}).pop(b, None)
        self.adj.get(b, {}).pop(a, None)

    @property
    def edges(self):
      seen, out = set(), []
      for u, nb in self.adj.items():
        for v in nb:
          if v not in seen:
            out.append((u, v))
        seen.add(u)
      return out

  def shortest_path(G, a, b):
    a, b = int(a), int(b)
    prev = {a: None}
    q = deque([a])
    while q:
      u = q.popleft()
      if u == b:
        break
      for v in G.adj[u]:
        if v not in prev:
          prev[v] = u
          q.append(v)
    path = [b]
    while path[-1] != a:
      path.append(prev[path[-1]])
    return path[::-1]

  stubs = {
    "networkx": types.SimpleNamespace(Graph=Graph, shortest_path=shortest_path),
    "fastremap": types.SimpleNamespace(unique=np.unique),
    "osteoid": types.SimpleNamespace(Skeleton=Skeleton, Bbox=object),
    "kimimaro": types.ModuleType("kimimaro"),
    "kimimaro.skeletontricks": ext,
  }
  stubs["kimimaro"].skeletontricks = ext
  saved = {k: sys.modules.get(k) for k in stubs}
  sys.modules.update(stubs)
  try:
    spec = importlib.util.spec_from_file_location("_reference_post", os.path.join(REF, "kimimaro", "post.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
  finally:
    for k, v in saved.items():
      if v is None:
        sys.modules.pop(k, None)
      else:
        sys.modules[k] = v
  return mod, Skeleton


def random_case(rng, loops):
  """A merged skeleton: 1-4 random trees on a jittered lattice (unique positions), `loops` extra edges inside trees."""
  verts, edges, radii = [], [], []
  n_parts = int(rng.integers(1, 5))
  for part in range(n_parts):
    n = int(rng.integers(2, 40))
    base = len(verts)
    origin = rng.uniform(0, 400, 3)
    pos = [origin]
    for i in range(1, n):
      p = int(rng.integers(max(0, i - 4), i))                 # elongated trees with side branches
      pos.append(pos[p] + rng.uniform(-1, 1, 3) * rng.choice([8.0, 30.0]) + np.array([20.0, 0, 0]))
      edges.append((base + p, base + i))
    extra = int(rng.integers(0, loops + 1))
    have = set(tuple(sorted(e)) for e in edges)
    for _ in range(extra):
      a, b = (base + int(v) for v in rng.integers(0, n, 2))
      if a != b and tuple(sorted((a, b))) not in have:
        have.add(tuple(sorted((a, b))))
        edges.append((a, b))
    verts += pos
    radii += list(rng.choice([2.0, 10.0, 40.0, 120.0], n))
  verts = np.round(np.array(verts), 2).astype(np.float32)
  order = rng.permutation(len(edges))
  edges = np.array(edges, dtype=np.uint32)[order]
  flip = rng.random(len(edges)) < 0.5
  edges[flip] = edges[flip][:, ::-1]
  return verts, edges, np.array(radii, dtype=np.float32)


NAMES = ("in", "loops", "dust", "join", "joinr", "ticks", "post")


def main():
  """One npz of five arrays: all vertices / radii / edges end to end, counts[case, NAMES index] = (vertices, edges) of
  each block in that order, params[case] = (tick threshold, dust threshold, join radius, loops knob)."""
  ref, Skeleton = load_reference_post()
  rng = np.random.default_rng(0xB200_9057)
  n_cases = 60
  V, R, E = [], [], []
  counts = np.zeros((n_cases, len(NAMES), 2), dtype=np.int64)
  params = np.zeros((n_cases, 4), dtype=np.float64)
  for c in range(n_cases):
    loops = 0 if c % 3 == 0 else 3
    v, e, r = random_case(rng, loops)
    tick = float(rng.choice([25.0, 60.0, 150.0, 400.0]))
    dust = float(rng.choice([0.0, 40.0, 200.0]))
    join_r = float(rng.choice([30.0, 150.0, np.inf]))
    params[c] = (tick, dust, join_r, loops)

    def fresh():
      return Skeleton(v.copy(), e.copy(), r.copy(), segid=c + 1).consolidate()

    out = {}
    out["loops"] = ref.remove_loops(fresh())
    out["dust"] = ref.remove_dust(fresh(), dust)
    out["join"] = ref.join_close_components(fresh(), radius=join_r)
    out["joinr"] = ref.join_close_components(fresh(), restrict_by_radius=True)
    out["ticks"] = ref.remove_ticks(ref.remove_loops(fresh()), tick)      # ticks are defined on trees
    out["post"] = ref.postprocess(Skeleton(v.copy(), e.copy(), r.copy(), segid=c + 1), dust, tick)
    for k, name in enumerate(NAMES):
      if name == "in":
        pv, pe, pr = v, e, r
      else:
        s = out[name].consolidate()
        pv, pe, pr = s.vertices.astype(np.float32), s.edges.astype(np.uint32), s.radii.astype(np.float32)
      counts[c, k] = (pv.shape[0], pe.shape[0])
      V.append(pv.reshape(-1, 3)), R.append(pr.reshape(-1)), E.append(pe.reshape(-1, 2))
  np.savez_compressed(OUT, vertices=np.concatenate(V), radii=np.concatenate(R), edges=np.concatenate(E),
                      counts=counts, params=params)
  print("wrote", OUT, os.path.getsize(OUT), "bytes,", n_cases, "cases")


if __name__ == "__main__":
  main()
