"""
Skeleton: the result type of skeletonize().

The reference returns osteoid.Skeleton objects (kimimaro/trace.py:34,182-192, intake.py:30,590;
cloud-volume re-exports the same class).  osteoid is not installed in this image, so this module
provides a compatible class with the fields and methods the reference's callers and tests touch
(SURVEY 8a row T): id, vertices (float32 Nx3), edges (uint32 Mx2), radii (float32), vertex_types
(uint8), space, transform (3x4 float32), extra_attributes; empty, from_path, simple_merge,
consolidate, merge, clone, cable_length, components, terminals, branches, voxel_space,
physical_space, to_swc / from_swc, to_precomputed / from_precomputed, equivalent.  The class is used unconditionally (HAVE_OSTEOID only records
whether the original is importable; converting is `osteoid.Skeleton(s.vertices, s.edges, s.radii, ...)`).
"""
import numpy as np

try:  # pragma: no cover - not available in this image
  import osteoid  # noqa: F401
  HAVE_OSTEOID = True
except Exception:  # ImportError and friends
  HAVE_OSTEOID = False


def _ident():
  return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]], dtype=np.float32)


class Skeleton:
  def __init__(self, vertices=None, edges=None, radii=None, vertex_types=None, segid=None,
               transform=None, space="voxel", extra_attributes=None):
    self.id = segid
    self.space = space
    self.vertices = (np.zeros((0, 3), np.float32) if vertices is None
                     else np.asarray(vertices, dtype=np.float32).reshape(-1, 3))
    self.edges = (np.zeros((0, 2), np.uint32) if edges is None
                  else np.asarray(edges, dtype=np.uint32).reshape(-1, 2))
    n = self.vertices.shape[0]
    self.radii = (-np.ones((n,), np.float32) if radii is None else np.asarray(radii, dtype=np.float32).reshape(-1))
    self.vertex_types = (np.zeros((n,), np.uint8) if vertex_types is None
                         else np.asarray(vertex_types, dtype=np.uint8).reshape(-1))
    self.transform = _ident() if transform is None else np.asarray(transform, dtype=np.float32).reshape(3, 4)
    self.extra_attributes = extra_attributes if extra_attributes is not None else [
      {"id": "radius", "data_type": "float32", "num_components": 1},
      {"id": "vertex_types", "data_type": "uint8", "num_components": 1},
    ]

  @classmethod
  def _from_arrays(cls, vertices, edges, radii, segid, transform, space):
    """Trusted fast path for the engine: arrays already have the right dtype and shape."""
    self = object.__new__(cls)
    self.id = segid
    self.space = space
    self.vertices = vertices
    self.edges = edges
    self.radii = radii
    self.vertex_types = np.zeros((vertices.shape[0],), np.uint8)
    self.transform = transform
    self.extra_attributes = [
      {"id": "radius", "data_type": "float32", "num_components": 1},
      {"id": "vertex_types", "data_type": "uint8", "num_components": 1},
    ]
    return self

  # -- basic -----------------------------------------------------------------------------------
  def empty(self):
    return self.vertices.size == 0 or self.edges.size == 0

  def __len__(self):
    return self.vertices.shape[0]

  def __repr__(self):
    return "Skeleton(segid={}, vertices=(shape={}), edges=(shape={}), radii=(shape={}), space='{}')".format(
      self.id, self.vertices.shape[0], self.edges.shape[0], self.radii.shape[0], self.space)

  def clone(self):
    return Skeleton(self.vertices.copy(), self.edges.copy(), self.radii.copy(), self.vertex_types.copy(),
                    segid=self.id, transform=self.transform.copy(), space=self.space,
                    extra_attributes=[dict(a) for a in self.extra_attributes])

  @classmethod
  def from_path(cls, path):
    path = np.asarray(path)
    n = path.shape[0]
    edges = np.zeros((max(n - 1, 0), 2), dtype=np.uint32)
    edges[:, 0] = np.arange(0, max(n - 1, 0))
    edges[:, 1] = np.arange(1, max(n, 1))
    return cls(path.astype(np.float32), edges)

  @classmethod
  def simple_merge(cls, skeletons):
    skeletons = list(skeletons)
    if len(skeletons) == 0:
      return cls()
    if type(skeletons[0]) is np.ndarray:
      skeletons = [skeletons]
    ct = 0
    edges = []
    for skel in skeletons:
      edges.append(skel.edges.astype(np.uint32) + np.uint32(ct))
      ct += skel.vertices.shape[0]
    first = skeletons[0]
    return cls(
      vertices=np.concatenate([s.vertices for s in skeletons], axis=0),
      edges=np.concatenate(edges, axis=0),
      radii=np.concatenate([s.radii for s in skeletons], axis=0),
      vertex_types=np.concatenate([s.vertex_types for s in skeletons], axis=0),
      segid=first.id, transform=first.transform, space=first.space,
    )

  def merge(self, skel):
    return Skeleton.simple_merge((self, skel)).consolidate()

  def consolidate(self, remove_disconnected_vertices=True):
    """Unique vertices (lexicographic x,y,z), edges remapped, each edge sorted, rows unique,
    self loops dropped, attributes of the first occurrence."""
    if self.vertices.shape[0] == 0:
      return Skeleton(segid=self.id, transform=self.transform, space=self.space)
    nodes, uniq_idx, inverse = np.unique(self.vertices, axis=0, return_index=True, return_inverse=True)
    inverse = inverse.reshape(-1)
    edges = inverse[self.edges.astype(np.int64)].reshape(-1, 2)
    edges = np.sort(edges, axis=1)
    if edges.shape[0]:
      edges = np.unique(edges, axis=0)
      edges = edges[edges[:, 0] != edges[:, 1]]
    skel = Skeleton(nodes, edges, self.radii[uniq_idx], self.vertex_types[uniq_idx], segid=self.id,
                    transform=self.transform, space=self.space, extra_attributes=self.extra_attributes)
    if remove_disconnected_vertices:
      skel = skel.remove_disconnected_vertices()
    return skel

  def remove_disconnected_vertices(self):
    used = np.zeros(self.vertices.shape[0], dtype=bool)
    used[self.edges.reshape(-1)] = True
    if used.all():
      return self
    remap = np.cumsum(used) - 1
    return Skeleton(self.vertices[used], remap[self.edges.astype(np.int64)], self.radii[used],
                    self.vertex_types[used], segid=self.id, transform=self.transform, space=self.space,
                    extra_attributes=self.extra_attributes)

  # -- geometry --------------------------------------------------------------------------------
  def cable_length(self):
    v1 = self.vertices[self.edges[:, 0]]
    v2 = self.vertices[self.edges[:, 1]]
    delta = (v2 - v1)
    delta *= delta
    return float(np.sum(np.sqrt(np.sum(delta, axis=1))))

  def _scale(self):
    return np.array([self.transform[0, 0], self.transform[1, 1], self.transform[2, 2]], dtype=np.float32)

  def voxel_space(self):
    if self.space == "voxel":
      return self.clone()
    skel = self.clone()
    skel.vertices = (skel.vertices - self.transform[:, 3]) / self._scale()
    skel.space = "voxel"
    return skel

  def physical_space(self):
    if self.space == "physical":
      return self.clone()
    skel = self.clone()
    skel.vertices = skel.vertices * self._scale() + self.transform[:, 3]
    skel.space = "physical"
    return skel

  # -- graph -----------------------------------------------------------------------------------
  def _degrees(self):
    deg = np.zeros(self.vertices.shape[0], dtype=np.int64)
    np.add.at(deg, self.edges.reshape(-1).astype(np.int64), 1)
    return deg

  def terminals(self):
    return np.flatnonzero(self._degrees() == 1)

  def branches(self):
    return np.flatnonzero(self._degrees() >= 3)

  def components(self):
    """Connected components as a list of Skeletons.  Like osteoid, the skeleton is consolidated first (duplicate
    vertices merged, isolated ones kept out of every component), so parts come in the order of their lowest
    consolidated vertex and keep that vertex order."""
    if self.vertices.shape[0] == 0:
      return []
    self = self.consolidate(remove_disconnected_vertices=False)
    n = self.vertices.shape[0]
    import scipy.sparse
    import scipy.sparse.csgraph
    e = self.edges.astype(np.int64)
    g = scipy.sparse.coo_matrix((np.ones(e.shape[0], np.int8), (e[:, 0], e[:, 1])), shape=(n, n))
    ncomp, lab = scipy.sparse.csgraph.connected_components(g, directed=False)
    out = []
    for c in range(ncomp):
      sel = lab == c
      if sel.sum() < 2:
        continue
      remap = np.cumsum(sel) - 1
      ee = e[sel[e[:, 0]]]
      out.append(Skeleton(self.vertices[sel], remap[ee], self.radii[sel], self.vertex_types[sel], segid=self.id,
                          transform=self.transform, space=self.space))
    return out

  @classmethod
  def equivalent(cls, first, second):
    if first.vertices.shape != second.vertices.shape or first.edges.shape != second.edges.shape:
      return False
    a, b = first.consolidate(), second.consolidate()
    return bool(np.array_equal(a.vertices, b.vertices) and np.array_equal(a.edges, b.edges))

  # -- SWC ---------------------------------------------------------------------------------------
  def to_swc(self, contributors=""):
    """Minimal SWC writer: one tree per connected component, root parent = -1."""
    n = self.vertices.shape[0]
    adj = [[] for _ in range(n)]
    for a, b in self.edges.tolist():
      adj[a].append(b)
      adj[b].append(a)
    lines = ["# generated by kimimaro_b200 (b200-teasar)", "# id type x y z radius parent"]
    seen = np.zeros(n, dtype=bool)
    ids = {}
    nxt = 1
    for root in range(n):
      if seen[root] or not adj[root]:
        continue
      stack = [(root, -1)]
      seen[root] = True
      while stack:
        v, parent = stack.pop()
        ids[v] = nxt
        x, y, z = self.vertices[v]
        lines.append("{} {} {:.6f} {:.6f} {:.6f} {:.6f} {}".format(
          nxt, int(self.vertex_types[v]), x, y, z, float(self.radii[v]), parent))
        me = nxt
        nxt += 1
        for u in adj[v]:
          if not seen[u]:
            seen[u] = True
            stack.append((u, me))
    return "\n".join(lines) + "\n"

  @classmethod
  def from_swc(cls, swcstr):
    verts, radii, types, parents, ids = [], [], [], [], {}
    for line in swcstr.splitlines():
      line = line.strip()
      if not line or line.startswith("#"):
        continue
      tok = line.split()
      ids[int(tok[0])] = len(verts)
      types.append(int(tok[1]))
      verts.append([float(tok[2]), float(tok[3]), float(tok[4])])
      radii.append(float(tok[5]))
      parents.append(int(tok[6]))
    edges = [[ids[p], i] for i, p in enumerate(parents) if p in ids]
    return cls(np.array(verts, np.float32).reshape(-1, 3), np.array(edges, np.uint32).reshape(-1, 2),
               np.array(radii, np.float32), np.array(types, np.uint8))

  # -- Neuroglancer precomputed (what Igneous stores per label after kimimaro.skeletonize) --------
  @property
  def radius(self):
    return self.radii

  def to_precomputed(self):
    """uint32 n_vertices, uint32 n_edges, float32 vertices [n][3], uint32 edges [m][2], then every entry of
    extra_attributes in order ([n][num_components] of its data_type); little endian, C order."""
    import struct
    verts = np.ascontiguousarray(self.vertices, dtype="<f4")
    edges = np.ascontiguousarray(self.edges, dtype="<u4")
    out = [struct.pack("<II", verts.shape[0], edges.shape[0]), verts.tobytes("C"), edges.tobytes("C")]
    for attr in self.extra_attributes:
      arr = np.asarray(getattr(self, attr["id"]))
      if arr.size != verts.shape[0] * int(attr["num_components"]):
        raise ValueError("attribute {} has {} values for {} vertices".format(attr["id"], arr.size, verts.shape[0]))
      out.append(np.ascontiguousarray(arr, dtype=np.dtype(attr["data_type"]).newbyteorder("<")).tobytes("C"))
    return b"".join(out)

  @classmethod
  def from_precomputed(cls, data, segid=None, vertex_attributes=None):
    """Inverse of to_precomputed; vertex_attributes defaults to radius (float32) + vertex_types (uint8) when the
    buffer is long enough for them, and to none for a bare vertices + edges buffer."""
    import struct
    if len(data) < 8:
      raise ValueError("precomputed skeleton buffer shorter than its 8-byte header")
    n, m = struct.unpack_from("<II", data, 0)
    off = 8
    need = off + 12 * n + 8 * m
    if len(data) < need:
      raise ValueError("precomputed skeleton buffer truncated: {} bytes, {} needed".format(len(data), need))
    verts = np.frombuffer(data, dtype="<f4", count=3 * n, offset=off).reshape(n, 3).copy()
    off += 12 * n
    edges = np.frombuffer(data, dtype="<u4", count=2 * m, offset=off).reshape(m, 2).copy()
    off += 8 * m
    if vertex_attributes is None:
      vertex_attributes = [] if len(data) == off else [
        {"id": "radius", "data_type": "float32", "num_components": 1},
        {"id": "vertex_types", "data_type": "uint8", "num_components": 1},
      ]
    skel = cls(verts, edges, segid=segid, extra_attributes=[dict(a) for a in vertex_attributes])
    for attr in vertex_attributes:
      dt = np.dtype(attr["data_type"]).newbyteorder("<")
      count = n * int(attr["num_components"])
      if len(data) < off + count * dt.itemsize:
        raise ValueError("precomputed skeleton buffer truncated in attribute " + attr["id"])
      arr = np.frombuffer(data, dtype=dt, count=count, offset=off).copy()
      off += count * dt.itemsize
      if int(attr["num_components"]) > 1:
        arr = arr.reshape(n, int(attr["num_components"]))
      setattr(skel, "radii" if attr["id"] == "radius" else attr["id"], arr)
    return skel
