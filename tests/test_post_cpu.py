"""
kimimaro_b200.post (chunk-stitch post-processing) on the CPU.

Pins, in the order of their strength:
  * the reference's OWN test cases, restated with their expected values: /root/reference/automated_test.py:335-382
    (find_cycle), :384-455 (join_close_components simple / complex / by_radius), :566-586 (remove_row),
    :611-629 (postprocess);
  * find_cycle and create_distance_graph against the reference's compiled extension (oracle/_ref, built from
    /root/reference/ext/skeletontricks by oracle/build_ref.py) on seeded random graphs -- skipped where it was never built;
  * hand-derived cases for the four loop rules and the tick rules of post.py:262-362, 446-563.
"""
import numpy as np
import pytest

from kimimaro_b200 import post
from kimimaro_b200.skeleton import Skeleton
import kimimaro_b200


# --- the reference's own cases --------------------------------------------------------------------------------------

def test_find_cycle_reference_cases():
  edges = np.array([[0, 1], [1, 2], [2, 0], [2, 3], [2, 4]], dtype=np.int32)
  assert np.array_equal(post.find_cycle(edges), [0, 2, 1, 0])

  edges = np.array([[0, 1], [1, 2], [2, 3], [3, 4], [4, 10], [10, 11], [11, 12], [12, 2], [4, 5], [5, 6], [6, 7]],
                   dtype=np.int32)
  assert np.array_equal(post.find_cycle(edges), [2, 12, 11, 10, 4, 3, 2])

  edges = np.array([[0, 1], [0, 20], [20, 21], [21, 22], [22, 23], [23, 21], [1, 2], [2, 3], [3, 4], [4, 5], [5, 6],
                    [6, 7], [7, 10], [10, 11], [11, 6]], dtype=np.int32)
  cycle = post.find_cycle(edges)
  assert np.array_equal(cycle, [21, 23, 22, 21]) or np.array_equal(cycle, [6, 11, 10, 7, 6])

  assert len(post.find_cycle(np.zeros((0, 2), np.int32))) == 0
  assert len(post.find_cycle(np.array([[0, 1], [1, 2], [2, 3]], np.int32))) == 0


def test_join_close_components_simple():
  skel = Skeleton([(0, 0, 0), (1, 0, 0), (10, 0, 0), (11, 0, 0)], edges=[(0, 1), (2, 3)], radii=[0, 1, 2, 3],
                  vertex_types=[0, 1, 2, 3], segid=1337)
  assert len(skel.components()) == 2

  res = kimimaro_b200.join_close_components(skel, radius=np.inf)
  assert len(res.components()) == 1

  res = kimimaro_b200.join_close_components(skel, radius=9)
  assert len(res.components()) == 1
  assert np.all(res.edges == [[0, 1], [1, 2], [2, 3]])
  assert res.id == 1337
  assert np.array_equal(res.radii, [0, 1, 2, 3])

  res = kimimaro_b200.join_close_components(skel, radius=8.5)
  assert len(res.components()) == 2

  with pytest.raises(ValueError):
    kimimaro_b200.join_close_components(skel, radius=0)
  assert kimimaro_b200.join_close_components([]).empty()
  one = Skeleton([(0, 0, 0), (1, 0, 0)], edges=[(0, 1)])
  assert Skeleton.equivalent(kimimaro_b200.join_close_components(one, radius=None), one)


def test_join_close_components_complex():
  skel = Skeleton([(0, 0, 0), (1, 0, 0), (4, 0, 0), (6, 0, 0), (20, 0, 0), (21, 0, 0), (0, 0, 5), (0, 0, 10)],
                  edges=[(0, 1), (2, 3), (4, 5), (6, 7)])
  assert len(skel.components()) == 4
  res = kimimaro_b200.join_close_components(skel, radius=np.inf)
  assert len(res.components()) == 1
  assert np.all(res.edges == [[0, 1], [0, 3], [1, 2], [3, 4], [4, 5], [5, 6], [6, 7]])
  # the same parts handed over as separate skeletons
  res2 = kimimaro_b200.join_close_components(skel.components(), radius=np.inf)
  assert Skeleton.equivalent(res, res2)


def test_join_close_components_by_radius():
  skel = Skeleton([(0, 0, 0), (1, 0, 0), (5, 0, 0), (11, 0, 0)], edges=[(0, 1), (2, 3)], radii=[100, 100, 100, 100],
                  vertex_types=[0, 1, 2, 3], segid=1337)
  res = kimimaro_b200.join_close_components(skel, restrict_by_radius=False)
  assert len(res.components()) == 1
  assert np.all(res.edges == [[0, 1], [1, 2], [2, 3]])

  res = kimimaro_b200.join_close_components(skel, restrict_by_radius=True)
  assert len(res.components()) == 1
  assert np.all(res.edges == [[0, 1], [1, 2], [2, 3]])

  skel.radii = np.array([1, 1, 1, 1], dtype=np.float32)
  res = kimimaro_b200.join_close_components(skel, restrict_by_radius=True)
  assert len(res.components()) == 2
  assert np.all(res.edges == [[0, 1], [2, 3]])

  skel.radii = np.array([1, 0.9, 3, 1], dtype=np.float32)
  res = kimimaro_b200.join_close_components(skel, restrict_by_radius=True)
  assert len(res.components()) == 2
  assert np.all(res.edges == [[0, 1], [2, 3]])

  skel.radii = np.array([1, 1, 3, 1], dtype=np.float32)
  res = kimimaro_b200.join_close_components(skel, restrict_by_radius=True)
  assert len(res.components()) == 1
  assert np.all(res.edges == [[0, 1], [1, 2], [2, 3]])


def test_remove_row():
  arr = np.array([[0, 1], [1, 2], [2, 1], [2, 2], [2, 3], [3, 4]])
  result = post.remove_row(arr, np.array([[1, 2]]))
  assert np.all(result == np.array([[0, 1], [2, 2], [2, 3], [3, 4]]))
  assert result.dtype == np.int32

  result = post.remove_row(np.array([[]]), np.array([[1, 2]]))
  assert np.all(result == np.array([]))
  assert result.size == 0


def test_postprocess_reference_case():
  skel = Skeleton([(0, 0, 0), (1, 0, 0), (4, 0, 0), (6, 0, 0), (20, 0, 0), (21, 0, 0), (0, 0, 5), (0, 0, 10)],
                  edges=[(0, 1), (2, 3), (4, 5), (6, 7), (0, 7), (1, 6)], segid=5)
  res = kimimaro_b200.postprocess(skel, dust_threshold=0, tick_threshold=0)
  ans = Skeleton([(4, 0, 0), (6, 0, 0), (20, 0, 0), (21, 0, 0)], edges=[(0, 1), (2, 3)])
  assert Skeleton.equivalent(res, ans)
  assert res.id == 5
  assert kimimaro_b200.postprocess(Skeleton(), 0, 0).empty()


# --- against the reference's compiled extension ---------------------------------------------------------------------

def _random_tree(rng, n):
  parent = [int(rng.integers(0, i)) for i in range(1, n)]
  edges = np.array([(p, i + 1) for i, p in enumerate(parent)], dtype=np.int64)
  return edges


def test_find_cycle_vs_reference_ext(ref_ext):
  if ref_ext is None or not hasattr(ref_ext, "find_cycle"):
    pytest.skip("oracle/_ref was never built here")
  rng = np.random.default_rng(0xC1C1E)
  for case in range(300):
    n = int(rng.integers(3, 60))
    edges = _random_tree(rng, n)
    extra = int(rng.integers(0, 4))
    rows = [tuple(e) for e in edges.tolist()]
    have = set(tuple(sorted(e)) for e in rows)
    for _ in range(extra):
      a, b = (int(v) for v in rng.integers(0, n, 2))
      if a != b and tuple(sorted((a, b))) not in have:
        have.add(tuple(sorted((a, b))))
        rows.append((a, b))
    rows = [rows[k] for k in rng.permutation(len(rows))]
    rows = [(b, a) if rng.random() < 0.5 else (a, b) for a, b in rows]
    relabel = rng.permutation(n + 5)[:n]                       # ids with gaps
    e32 = relabel[np.array(rows, dtype=np.int64)].astype(np.int32)
    want = np.asarray(ref_ext.find_cycle(e32.copy()))
    got = post.find_cycle(e32)
    assert np.array_equal(want, got), (case, e32.tolist(), want, got)


def test_create_distance_graph_vs_reference_ext(ref_ext):
  if ref_ext is None or not hasattr(ref_ext, "create_distance_graph"):
    pytest.skip("oracle/_ref was never built here")
  rng = np.random.default_rng(0xD157)
  for case in range(100):
    n = int(rng.integers(2, 80))
    edges = _random_tree(rng, n)
    edges = edges[rng.permutation(len(edges))]
    flip = rng.random(len(edges)) < 0.5
    edges[flip] = edges[flip][:, ::-1]
    skel = Skeleton(rng.uniform(0, 1000, (n, 3)).astype(np.float32), edges.astype(np.uint32))
    want = ref_ext.create_distance_graph(skel)
    got = post.create_distance_graph(skel)
    assert set(want.keys()) == set(got.keys()), case
    for k in want:
      assert abs(want[k] - got[k]) <= 1e-6 * max(1.0, abs(want[k])), (case, k, want[k], got[k])


def test_create_distance_graph_docstring_example():
  # skeletontricks.pyx:130-147: 1-2-3-4 with a 30 nm tick 2-5 and a 70 nm tick 3-6 (vertex 0 is unused)
  verts = [(0, 0, 0), (0, 0, 0), (60, 0, 0), (120, 0, 0), (180, 0, 0), (60, 30, 0), (120, 70, 0)]
  skel = Skeleton(verts, [(1, 2), (2, 3), (3, 4), (2, 5), (3, 6)])
  assert post.create_distance_graph(skel) == {(2, 1): 60.0, (3, 2): 60.0, (5, 2): 30.0, (4, 3): 60.0, (6, 3): 70.0}
  with pytest.raises(ValueError):
    post.create_distance_graph(Skeleton(verts, [(1, 2), (2, 3), (3, 1), (3, 4)]))


# --- loop rules -----------------------------------------------------------------------------------------------------

def _edge_set(skel):
  """Edges as pairs of COORDINATES (consolidate renumbers the vertices)."""
  v = [tuple(round(float(c), 3) for c in row) for row in skel.vertices.tolist()]
  return set(frozenset((v[a], v[b])) for a, b in skel.edges.tolist())


def _by_index(verts, pairs):
  v = [tuple(round(float(np.float32(c)), 3) for c in row) for row in verts]
  return set(frozenset((v[a], v[b])) for a, b in pairs)


def _ring(n, r=10.0):
  t = np.arange(n) * 2 * np.pi / n
  return [(float(r * np.cos(a)), float(r * np.sin(a)), 0.0) for a in t]


def test_remove_loops_isolated_ring_and_tree_untouched():
  ring = Skeleton(_ring(6), [(i, (i + 1) % 6) for i in range(6)])
  out = post.remove_loops(ring)
  assert out.edges.shape[0] == 0
  tree = Skeleton([(0, 0, 0), (1, 0, 0), (2, 0, 0), (1, 1, 0)], [(0, 1), (1, 2), (1, 3)])
  assert _edge_set(post.remove_loops(tree)) == _edge_set(tree)
  assert post.remove_loops(Skeleton()).empty()


def test_remove_loops_ring_on_a_stalk():
  # ring 0..5, stalk 0-6-7: the ring becomes one edge from the branch point 0 to the farthest ring node 3
  verts = _ring(6) + [(20.0, 0.0, 0.0), (30.0, 0.0, 0.0)]
  skel = Skeleton(verts, [(i, (i + 1) % 6) for i in range(6)] + [(0, 6), (6, 7)])
  out = post.remove_loops(skel)
  assert _edge_set(out) == _by_index(verts, [(0, 3), (0, 6), (6, 7)])
  assert len(post.find_cycle(out.edges.astype(np.int32))) == 0


def test_remove_loops_entrance_and_exit_keeps_the_short_arc():
  # ring 0..7 with tails at 0 and 2: the arc 0-1-2 (2 hops) stays, the 6-hop arc goes
  verts = _ring(8) + [(20.0, 0.0, 0.0), (0.0, 20.0, 0.0)]
  skel = Skeleton(verts, [(i, (i + 1) % 8) for i in range(8)] + [(0, 8), (2, 9)])
  out = post.remove_loops(skel)
  assert _edge_set(out) == _by_index(verts, [(0, 1), (1, 2), (0, 8), (2, 9)])


def test_remove_loops_many_gates_collapse_or_snip():
  # ring 0..5 with node 0 pulled to (1, 0, 0), tails at 0, 2, 4: the vertex nearest to the centroid of the three gates
  # (-3, 0, 0) is gate 0 itself, the farthest gate is 10.5 away from it
  verts = _ring(6) + [(20.0, 0.0, 0.0), (-10.0, 17.0, 0.0), (-10.0, -17.0, 0.0)]
  verts[0] = (1.0, 0.0, 0.0)
  edges = [(i, (i + 1) % 6) for i in range(6)] + [(0, 6), (2, 7), (4, 8)]
  wide = Skeleton(verts, edges, radii=np.full(9, 50.0, np.float32))
  out = post.remove_loops(wide)               # radius 50 >= 10.5: the ring collapses onto vertex 0
  assert _edge_set(out) == _by_index(verts, [(0, 2), (0, 4), (0, 6), (2, 7), (4, 8)])
  narrow = Skeleton(verts, edges, radii=np.full(9, 1.0, np.float32))
  out = post.remove_loops(narrow)             # radius 1 < 10.5: one snip (the walk's first edge) opens the ring
  assert len(post.find_cycle(out.edges.astype(np.int32))) == 0
  assert out.edges.shape[0] == len(edges) - 1
  assert len(out.components()) == 1


# --- tick rules -----------------------------------------------------------------------------------------------------

def _line(a, b, n):
  return [tuple(np.asarray(a, float) + (np.asarray(b, float) - np.asarray(a, float)) * k / n) for k in range(1, n + 1)]


def test_remove_ticks_threshold_and_shortest_first():
  # trunk 0..10 along x (length 100), a 30-long tick at x=50 and a 70-long branch at x=80
  verts = [(10.0 * k, 0.0, 0.0) for k in range(11)]
  edges = [(k, k + 1) for k in range(10)]
  tick = _line((50, 0, 0), (50, 30, 0), 3)
  branch = _line((80, 0, 0), (80, 70, 0), 7)
  n0 = len(verts)
  verts += tick
  edges += [(5, n0), (n0, n0 + 1), (n0 + 1, n0 + 2)]
  n1 = len(verts)
  verts += branch
  edges += [(8, n1)] + [(n1 + k, n1 + k + 1) for k in range(6)]
  skel = Skeleton(verts, edges)

  end = _by_index(verts, [(8, 9), (9, 10)])                               # the trunk's last 20
  tick_edges = _by_index(verts, [(5, n0), (n0, n0 + 1), (n0 + 1, n0 + 2)])  # the 30-long tick
  assert _edge_set(post.remove_ticks(skel, 0)) == _edge_set(skel)
  assert _edge_set(post.remove_ticks(skel, 15)) == _edge_set(skel)
  assert _edge_set(post.remove_ticks(skel, 20)) == _edge_set(skel)       # strictly below the threshold goes
  assert _edge_set(post.remove_ticks(skel, 25)) == _edge_set(skel) - end
  # at 40 the 20-long end goes first, vertex 8 dissolves (30 + 70 = 100 towards vertex 5), then the 30-long tick goes
  # and what is left is one unbranched path of 50 + 30 + 70
  for threshold in (40, 75, 1000):
    out = post.remove_ticks(skel, threshold)
    assert _edge_set(out) == _edge_set(skel) - end - tick_edges
    out = out.consolidate()
    assert len(out.components()) == 1
    assert len(out.terminals()) == 2
    assert abs(out.cable_length() - 150.0) < 1e-3

  # a single unbranched piece below the threshold is kept whole
  stub = Skeleton([(0, 0, 0), (1, 0, 0), (2, 0, 0)], [(0, 1), (1, 2)])
  assert _edge_set(post.remove_ticks(stub, 1000)) == _edge_set(stub)
  # a bare ring has no terminal: returned unchanged
  ring = Skeleton(_ring(5), [(i, (i + 1) % 5) for i in range(5)])
  assert _edge_set(post.remove_ticks(ring, 1000)) == _edge_set(ring)


def test_remove_dust():
  skel = Skeleton([(0, 0, 0), (10, 0, 0), (0, 5, 0), (1, 5, 0)], [(0, 1), (2, 3)])
  assert len(post.remove_dust(skel, 0).components()) == 2
  out = post.remove_dust(skel, 5)
  assert len(out.components()) == 1 and abs(out.cable_length() - 10) < 1e-6
  assert post.remove_dust(skel, 10).empty()          # strictly above the threshold survives


def test_postprocess_stitches_two_chunks():
  # two chunk skeletons of one neurite that overlap in one vertex, plus a small tick at the seam and a far crumb
  a = Skeleton([(0, 0, 0), (100, 0, 0), (200, 0, 0)], [(0, 1), (1, 2)], radii=[30, 30, 30], segid=7)
  b = Skeleton([(200, 0, 0), (300, 0, 0), (400, 0, 0), (200, 20, 0)], [(0, 1), (1, 2), (0, 3)], radii=[30, 30, 30, 5],
               segid=7)
  crumb = Skeleton([(5000, 0, 0), (5010, 0, 0)], [(0, 1)], radii=[5, 5], segid=7)
  merged = Skeleton.simple_merge([a, b, crumb])
  out = kimimaro_b200.postprocess(merged, dust_threshold=50, tick_threshold=100)
  assert out.id == 7
  assert len(out.components()) == 1
  assert np.array_equal(out.vertices, [(0, 0, 0), (100, 0, 0), (200, 0, 0), (300, 0, 0), (400, 0, 0)])
  assert np.array_equal(out.edges, [[0, 1], [1, 2], [2, 3], [3, 4]])


# --- goldens produced by running the reference's own post.py (scripts/make_post_golden.py) ---------------------------

def test_against_reference_run_goldens():
  import os
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "post_golden_v2.npz")
  z = np.load(path)
  names = ("in", "loops", "dust", "join", "joinr", "ticks", "post")       # scripts/make_post_golden.py: NAMES
  counts, params = z["counts"], z["params"]
  n = counts.shape[0]
  assert n >= 50 and counts.shape[1] == len(names)
  v_end = np.cumsum(counts[:, :, 0].reshape(-1))
  e_end = np.cumsum(counts[:, :, 1].reshape(-1))
  assert v_end[-1] == z["vertices"].shape[0] == z["radii"].shape[0] and e_end[-1] == z["edges"].shape[0]

  def block(c, name):
    k = c * len(names) + names.index(name)
    nv, ne = counts[c, names.index(name)]
    return (z["vertices"][v_end[k] - nv:v_end[k]], z["edges"][e_end[k] - ne:e_end[k]], z["radii"][v_end[k] - nv:v_end[k]])

  changed = {"loops": 0, "ticks": 0, "join": 0}
  for c in range(n):
    v, e, r = block(c, "in")
    tick, dust, join_r, _ = params[c].tolist()

    def fresh():
      return Skeleton(v.copy(), e.copy(), r.copy(), segid=c + 1).consolidate()

    got = {
      "loops": post.remove_loops(fresh()),
      "dust": post.remove_dust(fresh(), dust),
      "join": post.join_close_components(fresh(), radius=join_r),
      "joinr": post.join_close_components(fresh(), restrict_by_radius=True),
      "ticks": post.remove_ticks(post.remove_loops(fresh()), tick),
      "post": post.postprocess(Skeleton(v.copy(), e.copy(), r.copy(), segid=c + 1), dust, tick),
    }
    for name, sk in got.items():
      s = sk.consolidate()
      wv, we, wr = block(c, name)
      assert np.array_equal(s.vertices.reshape(-1, 3), wv), (c, name, "vertices")
      assert np.array_equal(s.edges.reshape(-1, 2), we), (c, name, "edges")
      assert np.array_equal(s.radii, wr), (c, name, "radii")
    base = fresh()
    changed["loops"] += int(got["loops"].edges.shape[0] != base.edges.shape[0])
    changed["ticks"] += int(got["ticks"].consolidate().edges.shape[0] != got["loops"].consolidate().edges.shape[0])
    changed["join"] += int(got["join"].consolidate().edges.shape[0] != base.edges.shape[0])
  # the cases must exercise the rules, not pass through them
  assert changed["loops"] >= 15 and changed["ticks"] >= 15 and changed["join"] >= 15, changed


def test_precomputed_round_trip():
  """The Neuroglancer precomputed skeleton encoding (what Igneous writes per label): header, vertices, edges, then the
  vertex attributes in order; checked against a hand-packed buffer and round-tripped."""
  import struct
  skel = Skeleton([(0, 0, 0), (1.5, 0, 0), (1.5, 2, 0)], [(0, 1), (1, 2)], radii=[1, 2, 3], vertex_types=[0, 5, 9], segid=3)
  buf = skel.to_precomputed()
  want = struct.pack("<II", 3, 2) + struct.pack("<9f", 0, 0, 0, 1.5, 0, 0, 1.5, 2, 0) + struct.pack("<4I", 0, 1, 1, 2)
  want += struct.pack("<3f", 1, 2, 3) + bytes([0, 5, 9])
  assert buf == want
  back = Skeleton.from_precomputed(buf, segid=3)
  assert back.id == 3
  assert np.array_equal(back.vertices, skel.vertices) and np.array_equal(back.edges, skel.edges)
  assert np.array_equal(back.radii, skel.radii) and np.array_equal(back.vertex_types, skel.vertex_types)
  bare = Skeleton.from_precomputed(buf[:8 + 36 + 16], vertex_attributes=[])
  assert np.array_equal(bare.edges, skel.edges) and np.all(bare.radii == -1)
  with pytest.raises(ValueError):
    Skeleton.from_precomputed(buf[:20])
  assert Skeleton.from_precomputed(Skeleton().to_precomputed()).empty()


def _trim_literal(skeleton, threshold):
  """The scan form of post.py:262-362 (min over the set, a list comprehension per dissolved branch point): what
  kimimaro_b200.post._trim_component must choose like, ties included."""
  from collections import defaultdict
  span = post.create_distance_graph(skeleton)
  edges = np.asarray(skeleton.edges).reshape(-1, 2)
  ids, deg = np.unique(edges, return_counts=True)
  tips = set(ids[deg == 1].tolist())
  arms = defaultdict(int)
  for v, c in zip(ids.tolist(), deg.tolist()):
    if c >= 3:
      arms[v] = c
  nbr = defaultdict(set)
  for a, b in edges.tolist():
    nbr[a].add(b)
    nbr[b].add(a)
  outer = set(e for e in span.keys() if (e[0] in tips or e[1] in tips))

  def dissolve(v):
    joined = [e for e in span.keys() if v in e]
    total = 0.0
    for e in joined:
      outer.discard(e)
      total += span[e]
      del span[e]
    ends = set(x for e in joined for x in e)
    ends.remove(v)
    span[tuple(ends)] = total
    outer.add(tuple(ends))
    arms[v] = 0

  while len(span) > 1 and outer:
    tick = min(outer, key=span.get)
    if span[tick] >= threshold:
      break
    a, b = tick
    path = post._hop_path(nbr, a, b)
    for u, v in zip(path[:-1], path[1:]):
      nbr[u].discard(v)
      nbr[v].discard(u)
    del span[tick]
    outer.remove(tick)
    arms[a] -= 1
    arms[b] -= 1
    if arms[a] == 2:
      dissolve(a)
    if arms[b] == 2:
      dissolve(b)
  return sorted((u, v) for u in nbr for v in nbr[u] if u < v)


def test_trim_component_equals_the_scan_form():
  rng = np.random.default_rng(0x71C5)
  removed = 0
  for case in range(40):
    n = int(rng.integers(5, 400))
    parent = [int(rng.integers(max(0, i - 6), i)) for i in range(1, n)]
    pos = np.zeros((n, 3), np.float32)
    for i in range(1, n):
      if case % 2 == 0:                       # lattice steps: many ticks of exactly equal length
        step = np.zeros(3)
        step[int(rng.integers(0, 3))] = float(rng.choice([-10, 10]))
      else:
        step = rng.uniform(-1, 1, 3) * 10 + np.array([5.0, 0, 0])
      pos[i] = pos[parent[i - 1]] + step
    skel = Skeleton(pos, np.array([(p, i + 1) for i, p in enumerate(parent)], np.uint32))
    for threshold in (15.0, 35.0, 80.0, 1e9):
      want = _trim_literal(skel.clone(), threshold)
      got = post._trim_component(skel.clone(), threshold)
      assert [tuple(e) for e in got.edges.tolist()] == want, (case, threshold)
      removed += int(len(want) < n - 1)
  assert removed >= 80
