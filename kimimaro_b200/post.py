"""
Chunk-stitch post-processing of merged skeletons (last item): postprocess = remove_dust ->
remove_loops -> join_close_components -> remove_ticks, the behaviour of /root/reference/kimimaro/post.py:49-222.

This is graph-sized host work in the reference (networkx / scipy / a small C++ helper) and it stays host work here
(numpy + scipy): a merged skeleton has 10^3..10^5 vertices, nothing a GPU is for, and it runs after the chunk merge,
outside the metric path.  Nothing here touches libb2t.so or the oracle.

What decides the RESULT follows the reference so that the same input gives the same skeleton:

  find_cycle             the depth-first walk of ext/skeletontricks/skeletontricks.hpp:206-297 (root = first entry of the
                         first edge, neighbours in first-seen order, last neighbour explored first).  Pinned against the
                         compiled reference extension in tests/test_post_cpu.py.
  create_distance_graph  skeletontricks.pyx:122-171 + skeletontricks.hpp:303-394: distances between critical points
                         (terminals and branch points) accumulated in float32 from the lowest-numbered terminal, keys
                         (larger id, smaller id).  Pinned against the compiled extension.
  remove_loops           post.py:446-563: the four cases by the number of branch points on the cycle.
  remove_ticks           post.py:262-362, including two quirks that decide results: a superedge made by fusing a former
                         branch point counts as a terminal superedge whatever its end points are (post.py:335), and ties
                         of the shortest superedge fall to Python's set order.
  join_close_components  post.py:89-218: pairs ranked by their nearest vertices, the fused part moves to the front of the
                         list, ties by the row-major first minimum of the distance matrix.

Not carried over: the dead branches of the reference (`np.all(radii_matrix) == np.inf`, post.py:179; the
`branch_counts == 1` stop, post.py:341, which no count can reach) and its IndexError on a component without a terminal in
remove_ticks (a bare ring is returned unchanged here).
"""
from collections import defaultdict, deque

import numpy as np

from .skeleton import Skeleton

__all__ = ["postprocess", "join_close_components", "remove_dust", "remove_loops", "remove_ticks", "find_cycle",
           "create_distance_graph", "remove_row", "path2edge"]

_NO_INDEX = np.iinfo(np.uint32).max


def postprocess(skeleton, dust_threshold=1500.0, tick_threshold=3000.0):
  """post.py:49-87.  Thresholds are physical lengths (the skeleton's units).  Returns a consolidated Skeleton."""
  label = skeleton.id
  skeleton = skeleton.consolidate()        # loops and ticks are read off a duplicate-free graph
  skeleton = remove_dust(skeleton, dust_threshold)
  skeleton = remove_loops(skeleton)
  skeleton = join_close_components(skeleton, restrict_by_radius=True)
  skeleton = remove_ticks(skeleton, tick_threshold)
  skeleton.id = label
  return skeleton.consolidate()


# ---------------------------------------------------------------------------------------------------------------------
# dust

def remove_dust(skeleton, dust_threshold):
  """Drop connected components whose cable length is not above dust_threshold (post.py:222-233)."""
  if skeleton.empty() or dust_threshold == 0:
    return skeleton
  return Skeleton.simple_merge([c for c in skeleton.components() if c.cable_length() > dust_threshold])


# ---------------------------------------------------------------------------------------------------------------------
# joining components

class _PairTable:
  """Nearest-vertex distance of every pair of parts: d[i, j] symmetric float32, at[i, j] = (vertex of i, vertex of j)
  for i < j only (post.py:134-163 fills the index of the upper triangle and leaves the lower one unset)."""

  def __init__(self, n):
    self.d = np.full((n, n), np.inf, dtype=np.float32)
    self.at = np.full((n, n, 2), _NO_INDEX, dtype=np.uint32)

  def keep(self, rows):
    """The table of [new part] + the parts listed in rows: the new part's row and column start out at infinity."""
    n = len(rows) + 1
    nxt = _PairTable(n)
    if len(rows):
      nxt.d[1:, 1:] = self.d[np.ix_(rows, rows)]
      nxt.at[1:, 1:] = self.at[np.ix_(rows, rows)]
    return nxt


def join_close_components(skeletons, radius=np.inf, restrict_by_radius=False):
  """
  Connect every component to its nearest other component through their two nearest vertices, nearest pair first, until
  one component is left or no pair is within `radius` (post.py:89-218).  restrict_by_radius: the reach becomes twice the
  largest radius and a pair is joined only if its gap is at most r1 + r2 of the two vertices.
  """
  if radius is None:
    radius = np.inf
  if radius <= 0:
    raise ValueError("radius must be greater than zero: " + str(radius))
  try:
    iter(skeletons)
  except TypeError:
    skeletons = [skeletons]

  parts = []
  for skeleton in skeletons:
    parts += skeleton.components()
  parts = [p.consolidate() for p in parts if not p.empty()]
  if len(parts) == 0:
    return Skeleton()
  if len(parts) == 1:
    return parts[0]

  from scipy.spatial import cKDTree

  if restrict_by_radius:
    radius = max(2 * np.max([np.max(p.radii) for p in parts]), 0)

  def rank(table, tree, i, j):
    a, b = parts[i], parts[j]
    gap, near = tree.query(b.vertices, k=1, distance_upper_bound=radius + 0.000001)   # the bound is exclusive
    vb = int(np.argmin(gap))
    va = int(near[vb])
    g = gap[vb]
    if restrict_by_radius and not np.isinf(g) and g > (a.radii[va] + b.radii[vb]):
      g = np.inf
    table.d[i, j] = table.d[j, i] = g
    table.at[i, j] = (min(va, _NO_INDEX), vb)

  table = _PairTable(len(parts))
  for i in range(len(parts)):
    tree = cKDTree(parts[i].vertices)
    for j in range(i + 1, len(parts)):
      rank(table, tree, i, j)

  while len(parts) > 1:
    best = np.min(table.d)
    if np.isinf(best) or best > radius:
      break
    i, j = np.unravel_index(np.argmin(table.d), table.d.shape)    # first minimum in row-major order: i < j
    a, b = parts[i], parts[j]
    fused = Skeleton.simple_merge([a, b])
    bridge = np.array([[table.at[i, j, 0], table.at[i, j, 1] + a.vertices.shape[0]]], dtype=np.uint32)
    fused.edges = np.concatenate([fused.edges, bridge])

    rest = [k for k in range(len(parts)) if k != i and k != j]
    parts = [fused] + [parts[k] for k in rest]
    table = table.keep(rest)
    tree = cKDTree(fused.vertices)
    for k in range(1, len(parts)):
      rank(table, tree, 0, k)

  return Skeleton.simple_merge(parts).consolidate()


# ---------------------------------------------------------------------------------------------------------------------
# loops

def find_cycle(edges):
  """
  One cycle of a connected graph as a closed node sequence [c, ..., c] (int32), or an empty array
  (skeletontricks.pyx:102-120, skeletontricks.hpp:206-297).  The walk is a depth-first search from edges[0, 0] over
  neighbour lists in first-seen order (duplicates dropped), taking the LAST listed neighbour first; it stops at the first
  node reached twice and returns the walk from that node's first visit on.
  """
  edges = np.asarray(edges)
  if edges.size == 0:
    return np.zeros((0,), dtype=np.int32)
  flat = edges.reshape(-1, 2).astype(np.int64)
  nbr = defaultdict(dict)                   # dict keys = insertion-ordered set
  for a, b in flat.tolist():
    nbr[a][b] = None
    nbr[b][a] = None

  todo = [(int(flat[0, 0]), -1, 0)]         # (node, the node it was reached from, depth)
  trail = []
  seen = set()
  node = -1
  while todo:
    node, came_from, depth = todo.pop()
    del trail[depth:]
    trail.append(node)
    if node in seen:
      break
    seen.add(node)
    for child in nbr[node]:
      if child != came_from:
        todo.append((child, node, depth + 1))

  if len(trail) <= 1:
    return np.zeros((0,), dtype=np.int32)
  start = len(trail) - 1
  for k in range(len(trail) - 1):
    if trail[k] == node:
      start = k
      break
  if len(trail) - start < 3:
    return np.zeros((0,), dtype=np.int32)
  return np.array(trail[start:], dtype=np.int32)


def path2edge(path):
  """A node sequence as its consecutive pairs (post.py:565-574)."""
  path = np.asarray(path)
  out = np.zeros((max(len(path) - 1, 0), 2), dtype=np.uint32)
  out[:, 0] = path[:-1]
  out[:, 1] = path[1:]
  return out


def _pair_key(pairs):
  pairs = np.asarray(pairs).astype(np.int64)
  return (pairs[:, 0] << 32) | pairs[:, 1]


def remove_row(array, rows2remove):
  """
  The edge list without the undirected edges of rows2remove; every row of the result is (low, high), int32
  (post.py:576-588: both arguments are ordered within each row first, every copy of a listed row goes).
  """
  array = np.asarray(array)
  if array.size == 0:
    return array.astype(np.int32, copy=False)
  array = np.sort(array.reshape(-1, 2), axis=1)
  drop = np.sort(np.asarray(rows2remove).reshape(-1, 2), axis=1)
  return array[~np.isin(_pair_key(array), _pair_key(drop))].astype(np.int32, copy=False)


def remove_loops(skeleton):
  """Break every cycle of every component (post.py:436-444); the rules are in _open_component."""
  if skeleton.empty():
    return skeleton
  parts = [_open_component(c) for c in skeleton.components()]
  return Skeleton.simple_merge(parts).consolidate(remove_disconnected_vertices=False)


def _open_component(skeleton):
  """
  post.py:446-563 for one connected component.  Until find_cycle finds nothing, by the number of cycle nodes that are
  branch points (degree >= 3) of the current graph:
    0  an isolated ring: all of its edges go;
    1  a ring on a stalk: the ring goes, one edge from the branch point to the ring node farthest from it is added;
    2  an entrance and an exit: the arc with more hops goes (half-and-half: the arc that does not wrap around the walk);
    3+ the ring collapses onto the skeleton vertex nearest to the centroid of its branch points, unless one of them is
       farther from that vertex than its radius -- then only the walk's first edge is cut.
  """
  xyz = skeleton.vertices
  edges = np.array(skeleton.edges, dtype=np.int64).reshape(-1, 2)

  while True:
    walk = find_cycle(edges.astype(np.int32))
    if len(walk) == 0:
      break
    ring_edges = np.sort(path2edge(walk).astype(np.int64), axis=1)      # in walk order
    ring = np.unique(ring_edges)
    ids, deg = np.unique(edges, return_counts=True)
    gates = ring[np.isin(ring, ids[deg >= 3])]                           # ascending ids

    if gates.shape[0] == 0:
      edges = remove_row(edges, ring_edges).astype(np.int64)
    elif gates.shape[0] == 1:
      away = np.sum((xyz[ring, :] - xyz[gates, :]) ** 2, 1)
      far = ring[np.argmax(away)]
      edges = remove_row(edges, ring_edges).astype(np.int64)
      edges = np.concatenate((edges, np.array([[gates[0], far]], dtype=np.int64)), 0)
    elif gates.shape[0] == 2:
      loop = np.array(walk[1:], dtype=np.int64)
      lo, hi = np.where(np.isin(loop, gates))[0][:2]
      if (hi - lo) < len(loop) / 2:
        kept = loop[lo:hi + 1]
      else:
        kept = np.concatenate((loop[hi:], loop[:lo + 1]), 0)
      kept_keys = _pair_key(np.sort(path2edge(kept).astype(np.int64), axis=1))
      edges = remove_row(edges, ring_edges[~np.isin(_pair_key(ring_edges), kept_keys)]).astype(np.int64)
    else:
      gate_xyz = xyz[gates, :]
      offset = xyz - np.mean(gate_xyz, axis=0)
      offset *= offset
      hub = int(np.argmin(np.sum(offset, axis=1)))
      reach = np.sqrt(np.max(np.sum((gate_xyz - xyz[hub, :]) ** 2, 1)))
      if reach > skeleton.radii[hub]:           # a wide ring must not pull everything to a point outside the neurite
        edges = remove_row(edges, ring_edges[:1, :]).astype(np.int64)
        continue
      edges = remove_row(edges, ring_edges).astype(np.int64)
      spokes = np.array([[g, hub] for g in gates.tolist() if g != hub], dtype=np.int64).reshape(-1, 2)
      edges = np.concatenate((edges, spokes), 0)

  out = skeleton.clone()
  out.edges = edges.astype(np.uint32)
  return out


# ---------------------------------------------------------------------------------------------------------------------
# ticks

def create_distance_graph(skeleton):
  """
  {(larger id, smaller id): float path length} between neighbouring critical points (terminals and branch points) of a
  single connected, cycle-free component (skeletontricks.pyx:122-171, skeletontricks.hpp:303-394).  Lengths accumulate in
  float32 along the walk away from the lowest-numbered terminal.  A cycle raises ValueError (the reference: RuntimeError
  out of the C++ helper); a component without a terminal gives {}.
  """
  xyz = np.asarray(skeleton.vertices, dtype=np.float32)
  edges = np.asarray(skeleton.edges).reshape(-1, 2).astype(np.int64)
  ids, deg = np.unique(edges, return_counts=True)
  tips = ids[deg == 1]
  if tips.shape[0] == 0:
    return {}
  critical = np.zeros(xyz.shape[0], dtype=bool)
  critical[tips] = True
  critical[ids[deg >= 3]] = True

  nbr = defaultdict(list)
  for a, b in edges.tolist():
    nbr[a].append(b)
    nbr[b].append(a)

  graph = {}
  seen = np.zeros(xyz.shape[0], dtype=bool)
  start = int(tips[0])
  todo = [(start, -1, np.float32(0.0), start)]    # (node, reached from, length since anchor, anchor)
  while todo:
    node, came_from, length, anchor = todo.pop()
    if seen[node]:
      raise ValueError("Cycle detected. Node: " + str(node))
    seen[node] = True
    if critical[node] and node != anchor:
      graph[(max(anchor, node), min(anchor, node))] = float(length)
      length = np.float32(0.0)
      anchor = node
    for child in nbr[node]:
      if child == came_from:
        continue
      d = xyz[node] - xyz[child]
      d = d * d
      step = np.sqrt(np.float32(np.float32(d[0] + d[1]) + d[2]))
      todo.append((child, node, np.float32(length + step), anchor))
  return graph


def remove_ticks(skeleton, threshold):
  """
  Remove terminal branches shorter than `threshold`, shortest first, re-ranking after every removal (post.py:235-260).
  """
  if skeleton.empty() or threshold == 0:
    return skeleton
  parts = [_trim_component(c, threshold) for c in skeleton.components()]
  return Skeleton.simple_merge(parts).consolidate(remove_disconnected_vertices=False)


def _hop_path(nbr, a, b):
  """Fewest-hop node path a..b over the adjacency sets (the unique path in a tree)."""
  prev = {a: None}
  queue = deque([a])
  while queue:
    u = queue.popleft()
    if u == b:
      break
    for v in nbr[u]:
      if v not in prev:
        prev[v] = u
        queue.append(v)
  if b not in prev:
    raise ValueError("no path between %d and %d" % (a, b))
  path = [b]
  while path[-1] != a:
    path.append(prev[path[-1]])
  return path[::-1]


def _trim_component(skeleton, threshold):
  """
  post.py:262-362 for one connected component, on the graph of distances between critical points.  While more than
  one superedge is left: take the shortest terminal superedge; stop if it is not below the threshold; delete its vertex
  path; a branch point that falls to two superedges is dissolved and its two superedges become one, which from then on
  ranks as a terminal superedge whatever its end points are (post.py:335).

  The reference scans every superedge per removal (min over a set, a list comprehension over the dict per dissolved
  branch point); the same choices are made here from a heap with lazy deletion and a node -> superedges index.  A tie of
  the minimum is the one case where the order of Python's set decides, and there the set is scanned like the reference.
  """
  import heapq
  if skeleton.empty():
    return skeleton
  span = create_distance_graph(skeleton)
  edges = np.asarray(skeleton.edges).reshape(-1, 2)
  ids, deg = np.unique(edges, return_counts=True)
  tips = set(ids[deg == 1].tolist())
  arms = defaultdict(int)                     # superedges left at a branch point; 0 for everything else
  for v, c in zip(ids.tolist(), deg.tolist()):
    if c >= 3:
      arms[v] = c

  nbr = defaultdict(set)
  for a, b in edges.tolist():
    nbr[a].add(b)
    nbr[b].add(a)

  outer = set(e for e in span.keys() if (e[0] in tips or e[1] in tips))
  born = {e: k for k, e in enumerate(span.keys())}          # the dict's insertion order, kept across deletions
  clock = len(born)
  at = defaultdict(set)                                       # node -> superedges that end there
  for e in span:
    at[e[0]].add(e)
    at[e[1]].add(e)
  heap = [(span[e], born[e], e) for e in outer]
  heapq.heapify(heap)

  def drop(e):
    del span[e]
    del born[e]
    for x in e:
      at[x].discard(e)

  def dissolve(v):
    nonlocal clock
    joined = sorted(at[v], key=born.get)                      # = [e for e in span if v in e]
    total = 0.0
    for e in joined:
      outer.discard(e)
      total += span[e]
    ends = set(x for e in joined for x in e)
    ends.remove(v)
    for e in joined:
      drop(e)
    fused = tuple(ends)
    if fused not in span:
      born[fused] = clock
      clock += 1
    span[fused] = total
    for x in fused:
      at[x].add(fused)
    outer.add(fused)
    heapq.heappush(heap, (total, born[fused], fused))
    arms[v] = 0

  def live(item):
    d, k, e = item
    return e in outer and born.get(e) == k and span[e] == d

  while len(span) > 1 and outer:
    while heap and not live(heap[0]):
      heapq.heappop(heap)
    first = heapq.heappop(heap)
    while heap and not live(heap[0]):
      heapq.heappop(heap)
    if heap and heap[0][0] == first[0]:
      tick = min(outer, key=span.get)                         # a tie: the set's own order decides (post.py:338)
      if tick != first[2]:
        heapq.heappush(heap, first)
    else:
      tick = first[2]
    if span[tick] >= threshold:
      break
    a, b = tick
    path = _hop_path(nbr, a, b) if len(nbr[a]) <= len(nbr[b]) else _hop_path(nbr, b, a)
    for u, v in zip(path[:-1], path[1:]):
      nbr[u].discard(v)
      nbr[v].discard(u)
    outer.remove(tick)
    drop(tick)
    arms[a] -= 1
    arms[b] -= 1
    if arms[a] == 2:
      dissolve(a)
    if arms[b] == 2:
      dissolve(b)

  left = sorted((u, v) for u in nbr for v in nbr[u] if u < v)
  out = skeleton.clone()
  out.edges = np.array(left, dtype=np.uint32).reshape(-1, 2)
  return out
