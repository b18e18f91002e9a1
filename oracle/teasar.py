"""
TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

numpy conductor of the CPU oracle: restates the control flow and dtype semantics of
  kimimaro/trace.py:36-356      (trace, compute_paths, find_soma_root, find_root, compute_pdrf)
  kimimaro/intake.py:58-221,434-593  (skeletonize, skeletonize_subset, compute_border_targets, merge)
  ext/skeletontricks/skeletontricks.pyx:490-760, 995-1045 (get_mapping, find_border_targets,
                                 compute_centroids, tie-breakers, CachedTargetFinder)
on top of the C restatements in oracle/oracle.c.  Skeletons are returned as plain dicts of
arrays {vertices f32[N,3], edges u32[M,2], radii f32[N]} so that the product's Skeleton class is
checked from outside (osteoid.Skeleton semantics restated: SURVEY A.8).

NumPy-2 scalar promotion rules are the target; numpy 2.3 is what this image has.
"""
from collections import defaultdict

import numpy as np

import oracle as orc

DEFAULT_TEASAR_PARAMS = {          # intake.py:47-56
  "scale": 1.5,
  "const": 300,
  "pdrf_scale": 100000,
  "pdrf_exponent": 4,
  "soma_acceptance_threshold": 3500,
  "soma_detection_threshold": 750,
  "soma_invalidation_const": 300,
  "soma_invalidation_scale": 2,
}

# The claim order of roll_invalidation_ball_inside_component that the oracle runs when none is asked for: the ENGINE's
# default ("window:1": rounds ordered by the reference's heap key in windows of one voxel).  "heap" is the reference's own
# order (== its compiled extension voxel for voxel; the engine's strict mode), "rounds" the hop-synchronous order of
# round 1 and "seq" the ordered process with canonical ties; see DESIGN.md 4.  tests/golden holds one set of vectors
# per engine order.
DEFAULT_INVALIDATION_MODE = "window:1"


class DimensionError(Exception):
  pass


# ------------------------------------------------------------------------------------------
# osteoid.Skeleton semantics on plain arrays
# ------------------------------------------------------------------------------------------
def skel_empty():
  return {"vertices": np.zeros((0, 3), np.float32), "edges": np.zeros((0, 2), np.uint32),
          "radii": np.zeros((0,), np.float32)}


def skel_from_path(path):
  path = np.asarray(path)
  n = path.shape[0]
  edges = np.zeros((max(n - 1, 0), 2), dtype=np.uint32)
  edges[:, 0] = np.arange(0, n - 1)
  edges[:, 1] = np.arange(1, n)
  return {"vertices": path.astype(np.float32), "edges": edges, "radii": -np.ones((n,), np.float32)}


def skel_simple_merge(skels):
  if len(skels) == 0:
    return skel_empty()
  verts, edges, radii = [], [], []
  off = 0
  for s in skels:
    verts.append(s["vertices"])
    edges.append(s["edges"].astype(np.uint32) + np.uint32(off))
    radii.append(s["radii"])
    off += s["vertices"].shape[0]
  return {"vertices": np.concatenate(verts).astype(np.float32).reshape(-1, 3),
          "edges": np.concatenate(edges).astype(np.uint32).reshape(-1, 2),
          "radii": np.concatenate(radii).astype(np.float32)}


def skel_consolidate(s):
  """unique vertices (lexicographic), edges remapped / each sorted / rows unique / self loops
  dropped, attributes of the first occurrence, vertices without edges removed."""
  if s["vertices"].shape[0] == 0:
    return skel_empty()
  nodes, uniq_idx, inverse = np.unique(s["vertices"], axis=0, return_index=True, return_inverse=True)
  inverse = inverse.reshape(-1)
  edges = inverse[s["edges"].astype(np.int64)].reshape(-1, 2)
  edges = np.sort(edges, axis=1)
  if edges.shape[0]:
    edges = np.unique(edges, axis=0)
    edges = edges[edges[:, 0] != edges[:, 1]]
  radii = s["radii"][uniq_idx]
  # remove_disconnected_vertices
  used = np.zeros(nodes.shape[0], dtype=bool)
  used[edges.reshape(-1)] = True
  if not used.all():
    remap = np.cumsum(used) - 1
    nodes = nodes[used]
    radii = radii[used]
    edges = remap[edges]
  return {"vertices": nodes.astype(np.float32), "edges": edges.astype(np.uint32).reshape(-1, 2),
          "radii": radii.astype(np.float32)}


# ------------------------------------------------------------------------------------------
# skeletontricks.pyx restatements (pinned against oracle/_ref in tests)
# ------------------------------------------------------------------------------------------
class CachedTargetFinder:
  """skeletontricks.pyx:995-1045 with canonical order T5 (DAF descending, ties by DESCENDING
  linear index == flip of a stable ascending sort; the reference's default argsort is unstable)."""
  def __init__(self, mask, daf):
    idx = np.flatnonzero(mask.ravel(order="F"))
    order = np.argsort(daf.ravel(order="F")[idx], kind="stable")[::-1]
    self.daf_indices = idx[order]
    self.cursor = 0

  def find_target(self, mask):
    flat = mask.ravel(order="F")
    d = self.daf_indices
    i = self.cursor
    n = d.size
    # chunked scan for the first still-valid entry
    while i < n:
      j = min(n, i + 4096)
      hit = np.flatnonzero(flat[d[i:j]])
      if hit.size:
        i += int(hit[0])
        self.cursor = i
        return tuple(int(v) for v in np.unravel_index(d[i], mask.shape, order="F"))
      i = j
    self.cursor = n
    return None


def first_label(labels):
  """skeletontricks.pyx:307-326: first non-zero in Fortran raster order."""
  flat = labels.ravel(order="F")
  nz = np.flatnonzero(flat)
  if nz.size == 0:
    return None
  return tuple(int(v) for v in np.unravel_index(nz[0], labels.shape, order="F"))


def get_mapping(orig_labels, cc_labels):
  """skeletontricks.pyx:490-525  { cc label: original label }.  Literally: in Fortran raster order, every voxel whose
  cc label differs from the previous voxel's writes remap[cc] = orig, so the LAST run start of a cc label decides.
  (Only matters when a cc label spans several original labels: engage_avocado_protection, intake.py:637.)"""
  cc = cc_labels.ravel(order="F")
  og = orig_labels.ravel(order="F")
  if cc.size == 0:
    return {}
  starts = np.concatenate(([0], np.flatnonzero(cc[1:] != cc[:-1]) + 1))
  rev = starts[::-1]
  _, first_in_rev = np.unique(cc[rev], return_index=True)
  return {int(cc[i]): int(og[i]) for i in rev[first_in_rev]}


def _f32(x):
  return np.float32(x)


def _distsq(p1x, p1y, p2x, p2y, wx, wy):
  a = _f32(wx * _f32(p1x - p2x))
  b = _f32(wy * _f32(p1y - p2y))
  return _f32(_f32(a * a) + _f32(b * b))


def _cornerness(x, y, sx, sy, wx, wy):
  # skeletontricks.pyx:732-748, including the reference's (sx-0.5) in the 4th corner's y
  h = np.float32(0.5)
  return min(_distsq(x, y, -h, -h, wx, wy), _distsq(x, y, _f32(sx - h), -h, wx, wy),
             _distsq(x, y, _f32(sx - h), _f32(sy - h), wx, wy), _distsq(x, y, -h, _f32(sx - h), wx, wy))


def _edgeness(x, y, sx, sy, wx, wy):
  # skeletontricks.pyx:716-730: double arithmetic (0.5 is a C double), rounded to float on return
  x, y, sx, sy, wx, wy = (float(v) for v in (x, y, sx, sy, wx, wy))
  return _f32(min(wx * (x - 0.5), wx * (sx - 0.5 - x), wy * (y - 0.5), wy * (sy - 0.5 - y)))


def compute_tiebreaker_maxima(px, py, x, y, centx, centy, sx, sy, wx, wy):
  """skeletontricks.pyx:650-715; all arguments are C floats there."""
  px, py, x, y, centx, centy, sx, sy, wx, wy = (_f32(v) for v in (px, py, x, y, centx, centy, sx, sy, wx, wy))
  cx = _f32(wx * sx / np.float32(2.0))
  cy = _f32(wy * sy / np.float32(2.0))
  d1 = _distsq(px, py, centx, centy, wx, wy)
  d2 = _distsq(x, y, centx, centy, wx, wy)
  if d2 < d1:
    return (x, y)
  elif d1 == d2:
    d1 = _distsq(px, py, cx, cy, wx, wy)
    d2 = _distsq(x, y, cx, cy, wx, wy)
    if d2 < d1:
      return (x, y)
    elif d1 == d2:
      d1 = _cornerness(px, py, sx, sy, wx, wy)
      d2 = _cornerness(x, y, sx, sy, wx, wy)
      if d2 < d1:
        return (x, y)
      elif d1 == d2:
        d1 = _edgeness(px, py, sx, sy, wx, wy)
        d2 = _edgeness(x, y, sx, sy, wx, wy)
        if d2 < d1:
          return (x, y)
  return (px, py)


def compute_centroids(labels, wx, wy):
  """skeletontricks.pyx:528-588: float32 running sums in x-outer / y-inner order."""
  wx = _f32(wx); wy = _f32(wy)
  sx, sy = labels.shape
  lab = np.ascontiguousarray(labels)          # C order == x outer, y inner
  flat = lab.ravel()
  xs = np.repeat(np.arange(sx, dtype=np.float32), sy)
  ys = np.tile(np.arange(sy, dtype=np.float32), sx)
  result = {}
  cx = _f32(_f32(wx * _f32(sx)) / _f32(2))
  cy = _f32(_f32(wy * _f32(sy)) / _f32(2))
  for label in np.unique(flat):
    if label == 0:
      continue
    sel = flat == label
    # sequential float32 accumulation (np.add.accumulate is a running sum, not pairwise)
    xsum = np.add.accumulate(xs[sel], dtype=np.float32)[-1]
    ysum = np.add.accumulate(ys[sel], dtype=np.float32)[-1]
    ct = _f32(np.count_nonzero(sel))
    px = _f32(_f32(wx * xsum) / ct)
    py = _f32(_f32(wy * ysum) / ct)
    if not (px - cx >= 0):
      px = _f32(px + wx)
    if not (py - cy >= 0):
      py = _f32(py + wy)
    result[int(label)] = (int(_f32(px / wx)), int(_f32(py / wy)))
  return result


def find_border_targets(dt, cc_labels, wx, wy):
  """skeletontricks.pyx:591-648.  Returns {label: (x, y)} with dict order = first raster encounter."""
  sx, sy = dt.shape
  centroids = None
  mx = defaultdict(float)
  pts = {}
  dtF = np.asfortranarray(dt)
  ccF = np.asfortranarray(cc_labels)
  flat_dt = dtF.ravel(order="F")
  flat_cc = ccF.ravel(order="F")
  sel = np.flatnonzero((flat_cc != 0) & (flat_dt != 0))
  if sel.size == 0:
    return pts
  labs = flat_cc[sel]
  vals = flat_dt[sel]
  # only voxels equal to their label's maximum can survive the scan; earlier smaller values are
  # overwritten, so folding over the maxima in raster order is the same computation.
  order_first = {}
  for l, i in zip(labs.tolist(), sel.tolist()):
    if l not in order_first:
      order_first[l] = i
  maxes = {}
  for l in order_first:
    maxes[l] = vals[labs == l].max()
  for l in order_first:                                  # dict insertion order (B.6)
    cand = sel[(labs == l) & (vals == maxes[l])]
    first = True
    for i in cand.tolist():
      x, y = i % sx, i // sx
      if first:
        pts[l] = (x, y); first = False
      else:
        if centroids is None:
          centroids = compute_centroids(cc_labels, wx, wy)
        px, py = pts[l]
        cx_, cy_ = centroids[l]
        r = compute_tiebreaker_maxima(px, py, x, y, cx_, cy_, sx, sy, wx, wy)
        pts[l] = (r[0], r[1])
  return pts


# ------------------------------------------------------------------------------------------
# kimimaro/trace.py
# ------------------------------------------------------------------------------------------
def is_power_of_two(num):
  if int(num) != num:
    return False
  num = int(num)
  return num != 0 and ((num & (num - 1)) == 0)


def compute_pdrf(dbf_max, pdrf_scale, pdrf_exponent, DBF, DAF, max_daf):
  """trace.py:315-356, same numpy expressions (NumPy-2 float32 scalar rules)."""
  f = lambda x: np.float32(x)
  with np.errstate(all="ignore"):
    M = f(1 / (dbf_max ** 1.01))
    PDRF = np.empty(DBF.shape, dtype=np.float32, order="F")
    np.multiply(DBF, M, out=PDRF)
    np.subtract(f(1), PDRF, out=PDRF)
    if is_power_of_two(pdrf_exponent) and (pdrf_exponent < (2 ** 16)):
      for _ in range(int(np.log2(pdrf_exponent))):
        PDRF *= PDRF
    else:
      np.power(PDRF, pdrf_exponent, out=PDRF)
    PDRF *= f(pdrf_scale)
    if max_daf != 0:
      DAF *= (1 / max_daf)
      PDRF += DAF
  return np.asfortranarray(PDRF)


def find_soma_root(DBF, dbf_max):
  """trace.py:269-289 (scipy.ndimage.center_of_mass of a boolean array == mean of coordinates, float64)."""
  maxima = (DBF == dbf_max)
  coords = np.vstack(np.where(maxima)).T
  com = np.asarray(coords.astype(np.float64).sum(axis=0) / float(coords.shape[0]), dtype=np.float32)
  root = np.argmin(np.sum((coords - com) ** 2, axis=1))
  return tuple(coords[root].astype(np.uint32))


def find_root(labels, anisotropy):
  any_voxel = first_label(labels)
  if any_voxel is None:
    return None
  _, target = orc.euclidean_distance_field(labels, any_voxel, anisotropy=anisotropy, return_max_location=True)
  return target


def trace(labels, DBF, scale=10, const=10, anisotropy=(1, 1, 1),
          soma_detection_threshold=1100, soma_acceptance_threshold=4000,
          pdrf_scale=5000, pdrf_exponent=16, soma_invalidation_scale=0.5, soma_invalidation_const=0,
          fix_branching=True, manual_targets_before=None, manual_targets_after=None, root=None,
          max_paths=None, voxel_graph=None, invalidation_mode=None, return_paths=False):
  """trace.py:36-194."""
  assert voxel_graph is None, "voxel_graph is out of scope (SURVEY 8f N4)"
  invalidation_mode = invalidation_mode or DEFAULT_INVALIDATION_MODE
  manual_targets_before = [] if manual_targets_before is None else manual_targets_before
  manual_targets_after = [] if manual_targets_after is None else manual_targets_after
  dbf_max = np.max(DBF)
  labels = np.asfortranarray(labels).view(np.uint8) if labels.dtype == bool else np.asfortranarray(labels, dtype=np.uint8)
  labels = labels.copy(order="F")
  DBF = np.array(DBF, dtype=np.float32, order="F")

  soma_mode = False
  if dbf_max > soma_detection_threshold:
    labels, num_voxels_filled = orc.fill_voids(labels)
    if num_voxels_filled > 0:
      DBF = orc.edt(labels, anisotropy=anisotropy, black_border=bool(np.all(labels)))
    dbf_max = np.max(DBF)
    soma_mode = dbf_max > soma_acceptance_threshold

  soma_radius = 0.0
  if soma_mode:
    if root is not None:
      manual_targets_before.insert(0, root)
    root = find_soma_root(DBF, dbf_max)
    soma_radius = dbf_max * soma_invalidation_scale + soma_invalidation_const
  elif root is None:
    root = find_root(labels, anisotropy)

  if root is None:
    return skel_empty() if not return_paths else (skel_empty(), [])
  root = tuple(root)

  free_space_radius = 0 if not soma_mode else DBF[root]
  DBF[DBF == 0] = np.inf                                    # zero2inf
  DAF, target = orc.euclidean_distance_field(labels, root, anisotropy=anisotropy,
                                             free_space_radius=free_space_radius, return_max_location=True)
  DAF[DAF == np.inf] = 0                                    # inf2zero
  target_finder = CachedTargetFinder(labels, DAF)
  PDRF = compute_pdrf(dbf_max, pdrf_scale, pdrf_exponent, DBF, DAF, DAF[target])
  del DAF

  if not fix_branching:
    parents = orc.parental_field(PDRF, root)
  else:
    parents = PDRF

  if soma_mode:
    _, labels = orc.roll_invalidation_ball_inside_component(
      labels, DBF, soma_invalidation_scale, soma_invalidation_const, anisotropy, [root], mode=invalidation_mode)
  elif len(manual_targets_before) == 0:
    manual_targets_before.append(target)

  paths = compute_paths(root, labels, DBF, target_finder, parents, scale, const, anisotropy,
                        soma_mode, soma_radius, fix_branching, manual_targets_before, manual_targets_after,
                        max_paths, invalidation_mode)

  skel = skel_consolidate(skel_simple_merge([skel_from_path(p) for p in paths if len(p) > 0]))
  verts = skel["vertices"].flatten().astype(np.uint32)
  skel["radii"] = DBF[verts[::3], verts[1::3], verts[2::3]].astype(np.float32)
  if return_paths:
    return skel, paths
  return skel


def compute_paths(root, labels, DBF, target_finder, parents, scale, const, anisotropy, soma_mode, soma_radius,
                  fix_branching, manual_targets_before, manual_targets_after, max_paths, invalidation_mode):
  """trace.py:196-267."""
  paths = []
  valid_labels = int(np.count_nonzero(labels))
  if max_paths is None:
    max_paths = valid_labels
  if len(manual_targets_before) + len(manual_targets_after) >= max_paths:
    return []
  parents[tuple(root)] = 0
  anisotropy = np.asarray(anisotropy, dtype=np.float32)
  while (valid_labels > 0 or manual_targets_before or manual_targets_after) and len(paths) < max_paths:
    if manual_targets_before:
      target = manual_targets_before.pop()
    elif valid_labels == 0:
      target = manual_targets_after.pop()
    else:
      target = target_finder.find_target(labels)
    if fix_branching:
      path = orc.railroad(parents, target)
    else:
      path = orc.path_from_parents(parents, target)
    if soma_mode:
      # trace.py:246-251 with its dtypes: uint32 path minus uint32 root wraps
      with np.errstate(over="ignore"):
        dist_to_soma_root = np.linalg.norm(anisotropy * (path - np.asarray(root, dtype=np.uint32)), axis=1)
      path = np.concatenate((path[:1, :], path[dist_to_soma_root > soma_radius, :]))
    if valid_labels > 0:
      invalidated, labels = orc.roll_invalidation_ball_inside_component(
        labels, DBF, scale, const, anisotropy, path, mode=invalidation_mode)
      valid_labels -= invalidated
    if fix_branching:
      parents[path[:, 0], path[:, 1], path[:, 2]] = 0.0
    paths.append(path)
  return paths


# ------------------------------------------------------------------------------------------
# kimimaro/intake.py
# ------------------------------------------------------------------------------------------
_POOL_FN = None


def _pool_call(segid):
  return _POOL_FN(segid)


def format_labels(labels):
  labels = np.copy(labels, order="F")
  if labels.dtype == bool:
    labels = labels.view(np.uint8)
  original_shape = labels.shape
  while labels.ndim < 3:
    labels = labels[..., np.newaxis]
  while labels.ndim > 3:
    if labels.shape[-1] == 1:
      labels = labels[..., 0]
    else:
      raise DimensionError(
        "Input labels may be no more than three non-trivial dimensions. Got: {}".format(original_shape))
  return labels


def find_objects(cc_labels, n):
  """utility.py:85-102: list of per-label slice triples (None if absent)."""
  import scipy.ndimage
  all_slices = scipy.ndimage.find_objects(cc_labels.T, max_label=n)
  return [(slcs and slcs[::-1]) for slcs in all_slices]


def compute_border_targets(cc_labels, anisotropy):
  """intake.py:544-585."""
  sx, sy, sz = cc_labels.shape
  planes = (
    (cc_labels[:, :, 0], (0, 1), lambda x, y: (x, y, 0)),
    (cc_labels[:, :, -1], (0, 1), lambda x, y: (x, y, sz - 1)),
    (cc_labels[:, 0, :], (0, 2), lambda x, z: (x, 0, z)),
    (cc_labels[:, -1, :], (0, 2), lambda x, z: (x, sy - 1, z)),
    (cc_labels[0, :, :], (1, 2), lambda y, z: (0, y, z)),
    (cc_labels[-1, :, :], (1, 2), lambda y, z: (sx - 1, y, z)),
  )
  target_list = defaultdict(set)
  for plane, dims, rotatefn in planes:
    wx, wy = anisotropy[dims[0]], anisotropy[dims[1]]
    plane = np.copy(plane, order="F")
    cc_plane, _ = orc.connected_components(plane)
    dt_plane = orc.edt(cc_plane, black_border=True, anisotropy=(wx, wy))
    plane_targets = find_border_targets(dt_plane, cc_plane, wx, wy)
    remapping = get_mapping(plane[..., np.newaxis], cc_plane[..., np.newaxis])
    for label, pt in plane_targets.items():
      label = remapping[label]
      target_list[label].add(rotatefn(int(pt[0]), int(pt[1])))
  out = defaultdict(lambda: np.array([], np.uint32))
  for label, pts in target_list.items():
    out[label] = np.array(list(pts), dtype=np.uint32)
  return out


def find_avocado_fruit(labels, cx, cy, cz, background=0):
  """skeletontricks.pyx:905-992: walk the six axis rays from (cx,cy,cz); the first foreign label met on each ray
  (a ray that meets background first says nothing) votes for the fruit around the pit.  The rays towards smaller
  coordinates stop BEFORE index 0, like the reference's range(c, 0, -1)."""
  sx, sy, sz = labels.shape[:3]
  if cx >= sx or cy >= sy or cz >= sz:
    raise ValueError("<{},{},{}> must be be contained within shape <{},{},{}>".format(cx, cy, cz, sx, sy, sz))
  label = labels[cx, cy, cz]
  changes = []

  def ray(line, rng):
    for i in rng:
      if line[i] == background:
        return
      if line[i] != label:
        changes.append(line[i])
        return
  ray(labels[:, cy, cz], range(cx, sx))
  ray(labels[:, cy, cz], range(cx, 0, -1))
  ray(labels[cx, :, cz], range(cy, sy))
  ray(labels[cx, :, cz], range(cy, 0, -1))
  ray(labels[cx, cy, :], range(cz, sz))
  ray(labels[cx, cy, :], range(cz, 0, -1))
  if len(changes) < 3:                       # too little info to make a decision
    return (label, label)
  allowed_differences = 1 if len(changes) > 3 else 0
  uniq, cts = np.unique(changes, return_counts=True)
  candidate_fruit_index = np.argmax(cts)
  differences = len(changes) - cts[candidate_fruit_index]
  if differences > allowed_differences:      # lots of labels around the candidate pit: not an avocado
    return (label, label)
  return (label, uniq[candidate_fruit_index])


def fill_voids_2d(plane):
  """fill_voids.fill on a 2-D image (intake.py:671-676): background not 4-connected to the image border is filled."""
  import scipy.ndimage
  if plane.size == 0:
    return plane
  return scipy.ndimage.binary_fill_holes(plane)


def renumber(cc_labels):
  """fastremap.renumber(arr, in_place=True) (intake.py:636): 1..N in order of first appearance in memory order
  (Fortran raster for the F-ordered cc_labels), 0 stays 0.  Returns (renumbered, {old: new})."""
  flat = cc_labels.ravel(order="F")
  uniq, first = np.unique(flat, return_index=True)
  order = np.argsort(first, kind="stable")
  mapping = {}
  nxt = 1
  for u in uniq[order]:
    if u == 0:
      mapping[0] = 0
    else:
      mapping[int(u)] = nxt
      nxt += 1
  lut = np.zeros(int(uniq.max()) + 1, dtype=cc_labels.dtype)
  for k, v in mapping.items():
    lut[k] = v
  return lut[cc_labels], mapping


def _intset(values):
  """set(fastremap.unique(...)): ints hash like numpy integers, so a set built from the same values in the same
  (sorted) insertion order iterates in the same order as the reference's."""
  return set(int(v) for v in values)


def engage_avocado_protection_single_pass(cc_labels, all_dbf, n_cc, candidates):
  """intake.py:646-704."""
  candidates = [label for label in candidates if label != 0]
  unchanged, changed = set(), set()
  if len(candidates) == 0:
    return cc_labels, unchanged, changed

  def paint_walls(binimg):
    binimg[:, :, 0] = fill_voids_2d(binimg[:, :, 0])
    binimg[:, :, -1] = fill_voids_2d(binimg[:, :, -1])
    binimg[:, 0, :] = fill_voids_2d(binimg[:, 0, :])
    binimg[:, -1, :] = fill_voids_2d(binimg[:, -1, :])
    binimg[0, :, :] = fill_voids_2d(binimg[0, :, :])
    binimg[-1, :, :] = fill_voids_2d(binimg[-1, :, :])
    return binimg

  slcs = find_objects(cc_labels, n_cc)
  for label in candidates:
    slc = slcs[label - 1]
    offset = np.array([s.start for s in slc])
    binimg = paint_walls(np.asfortranarray(cc_labels[slc] == label))    # image of the pit
    prod = binimg * all_dbf[slc]
    coord = np.array(np.unravel_index(np.argmax(prod.ravel(order="F")), prod.shape, order="F")) + offset
    pit, fruit = find_avocado_fruit(cc_labels, int(coord[0]), int(coord[1]), int(coord[2]))
    pit, fruit = int(pit), int(fruit)
    if pit == fruit and pit not in changed:
      unchanged.add(pit)
    else:
      unchanged.discard(pit)
      unchanged.discard(fruit)
      changed.add(pit)
      changed.add(fruit)
      binimg |= (cc_labels[slc] == fruit)
    binimg, N = orc.fill_voids(binimg)
    cc_labels[slc] = cc_labels[slc] * ~binimg + np.asarray(fruit, cc_labels.dtype) * binimg
  return cc_labels, unchanged, changed


def engage_avocado_protection(cc_labels, all_dbf, n_cc, remapping, soma_detection_threshold, edtfn):
  """intake.py:600-644: a nucleus segmented apart from its cell (the pit of an avocado) is given the label of the
  fruit around it; up to 20 passes for nested cases; then renumber and re-derive the cc -> original label map.
  Returns (cc_labels, all_dbf, remapping, n_cc)."""
  orig_cc_labels = np.copy(cc_labels, order="F")
  unchanged = set()
  for _ in range(20):
    candidates = _intset(np.unique(cc_labels * (all_dbf > soma_detection_threshold / 2.5)))
    candidates -= unchanged
    candidates.discard(0)
    cc_labels, unchanged_this_cycle, changes = engage_avocado_protection_single_pass(
      cc_labels, all_dbf, n_cc, candidates)
    unchanged |= unchanged_this_cycle
    if len(changes) == 0:
      break
    all_dbf = edtfn(cc_labels)
  cc_labels, _ = renumber(cc_labels)
  cc_labels = np.asfortranarray(cc_labels)
  cc_remapping = get_mapping(orig_cc_labels, cc_labels)
  adjusted_remapping = {}
  for new_cc, cc in cc_remapping.items():
    if cc in remapping:
      adjusted_remapping[new_cc] = remapping[cc]
  return cc_labels, all_dbf, adjusted_remapping, int(cc_labels.max())


def fill_all_holes(cc_labels, n_cc, return_fill_count=False):
  """intake.py:747-794: fill the holes of every connected component, in label order; a component that is swallowed
  by an earlier one disappears (and is not visited any more).  In place on the F-ordered cc_labels."""
  labels = np.unique(cc_labels)
  labels_set = set(int(l) for l in labels)
  labels_set.discard(0)
  all_slices = find_objects(cc_labels, n_cc)
  pixels_filled = 0
  for label in labels:
    label = int(label)
    if label not in labels_set:
      continue
    slices = all_slices[label - 1]
    if slices is None:
      continue
    binary_image = np.asfortranarray(cc_labels[slices] == label)
    binary_image, N = orc.fill_voids(binary_image)
    pixels_filled += N
    if N == 0:
      continue
    sub_labels = set(int(l) for l in np.unique(cc_labels[slices] * binary_image))
    sub_labels.remove(label)
    labels_set -= sub_labels
    cc_labels[slices] = cc_labels[slices] * ~binary_image + np.asarray(label, cc_labels.dtype) * binary_image
  if return_fill_count:
    return cc_labels, pixels_filled
  return cc_labels


def point_to_point(binary_img, start, end, anisotropy=(1, 1, 1), pdrf_scale=100000, pdrf_exponent=4):
  """trace.py:358-390.  dijkstra3d.dijkstra(PDRF, end, start) = the parent walk from `start` through the node-weighted
  field grown from `end` (parental_field + path_from_parents: the path reads source -> target, i.e. end first)."""
  binary_img = np.asfortranarray(binary_img).view(np.uint8) if binary_img.dtype == bool else np.asfortranarray(binary_img, dtype=np.uint8)
  DBF = orc.edt(binary_img, anisotropy=anisotropy, black_border=True)
  dbf_max = np.max(DBF)
  DBF[DBF == 0] = np.inf                                    # zero2inf
  DAF, target = orc.euclidean_distance_field(binary_img, tuple(start), anisotropy=anisotropy, return_max_location=True)
  DAF[DAF == np.inf] = 0                                    # inf2zero
  PDRF = compute_pdrf(dbf_max, pdrf_scale, pdrf_exponent, DBF, DAF, DAF[tuple(target)])
  parents = orc.parental_field(PDRF, tuple(end))
  path = orc.path_from_parents(parents, tuple(start))
  skel = skel_from_path(path)
  verts = skel["vertices"].flatten().astype(np.uint32)
  skel["radii"] = DBF[verts[::3], verts[1::3], verts[2::3]].astype(np.float32)
  return skel


def connect_points(labels, start, end, anisotropy=(1, 1, 1), fill_holes=False, in_place=False, pdrf_scale=100000,
                   pdrf_exponent=4):
  """intake.py:268-313."""
  anisotropy = np.array(anisotropy, dtype=np.float32)
  start, end = tuple(start), tuple(end)
  labels = format_labels(np.asarray(labels).astype(bool))
  start, end = start + (0,) * (3 - len(start)), end + (0,) * (3 - len(end))
  cc_labels, _ = orc.connected_components(labels.view(np.uint8))
  if cc_labels[start] == 0 or cc_labels[start] != cc_labels[end]:
    raise ValueError("Cannot extract centerline from disconnected components.")
  skel = point_to_point(labels, start, end, anisotropy=tuple(float(a) for a in anisotropy), pdrf_scale=pdrf_scale,
                        pdrf_exponent=pdrf_exponent)
  skel["vertices"] = skel["vertices"] * anisotropy
  return skel


def skeletonize(all_labels, teasar_params=DEFAULT_TEASAR_PARAMS, anisotropy=(1, 1, 1), object_ids=None,
                dust_threshold=1000, fix_branching=True, fix_borders=True,
                extra_targets_before=(), extra_targets_after=(), invalidation_mode=None,
                only_cc=None, timings=None, parallel=1, fill_holes=False, fix_avocados=False):
  """intake.py:58-221 + 434-517 (parallel==1 path).  Returns {orig id: skeleton dict}."""
  import time
  invalidation_mode = invalidation_mode or DEFAULT_INVALIDATION_MODE
  t0 = time.time()
  anisotropy = np.array(anisotropy, dtype=np.float32)
  all_labels = format_labels(all_labels)
  if object_ids is not None:
    all_labels = all_labels * np.isin(all_labels, object_ids)
  if all_labels.size <= dust_threshold:
    return {}
  minlabel, maxlabel = all_labels.min(), all_labels.max()
  if minlabel == 0 and maxlabel == 0:
    return {}
  cc_labels, n_cc = orc.connected_components(all_labels)
  remapping = get_mapping(all_labels, cc_labels)
  if fill_holes:
    cc_labels = fill_all_holes(cc_labels, n_cc)             # intake.py:166-167
  def points_to_labels(pts):
    mapping = defaultdict(list)
    for pt in pts:
      pt = tuple(pt)
      mapping[int(cc_labels[pt])].append(pt)
    return mapping
  extra_targets_before = points_to_labels(extra_targets_before)
  extra_targets_after = points_to_labels(extra_targets_after)
  edtfn = lambda labels: orc.edt(labels, anisotropy=anisotropy, black_border=bool(minlabel == maxlabel))
  all_dbf = edtfn(cc_labels)
  if fix_avocados:                                            # intake.py:187-193
    cc_labels, all_dbf, remapping, n_cc = engage_avocado_protection(
      cc_labels, all_dbf, n_cc, remapping, teasar_params.get("soma_detection_threshold", 0), edtfn)
  counts = np.bincount(cc_labels.ravel(order="K"), minlength=n_cc + 1)
  cc_segids = [sid for sid in range(1, n_cc + 1) if counts[sid] > dust_threshold]
  all_slices = find_objects(cc_labels, n_cc)
  border_targets = defaultdict(list)
  if fix_borders:
    border_targets = compute_border_targets(cc_labels, anisotropy)
  t1 = time.time()

  def trace_one(segid):
    slices = all_slices[segid - 1]
    if slices is None:
      return None
    minpt = np.array([s.start for s in slices])
    vol = np.prod([s.stop - s.start for s in slices])
    if vol <= 1:
      return None
    labels = (cc_labels[slices] == segid)
    dbf = np.where(labels, all_dbf[slices], 0.0).astype(np.float32)
    manual_targets_before, manual_targets_after, root = [], [], None
    def translate_to_roi(targets):
      targets = np.array(targets)
      targets -= minpt.astype(np.uint32)
      return targets.tolist()
    if len(border_targets[segid]) > 0:
      manual_targets_before = translate_to_roi(border_targets[segid])
      root = manual_targets_before.pop()
    if segid in extra_targets_before and len(extra_targets_before[segid]) > 0:
      manual_targets_before.extend(translate_to_roi(extra_targets_before[segid]))
    if segid in extra_targets_after and len(extra_targets_after[segid]) > 0:
      manual_targets_after.extend(translate_to_roi(extra_targets_after[segid]))
    skel = trace(labels, dbf, anisotropy=anisotropy, fix_branching=fix_branching,
                 manual_targets_before=manual_targets_before, manual_targets_after=manual_targets_after,
                 root=root, invalidation_mode=invalidation_mode, **teasar_params)
    if skel["vertices"].shape[0] == 0:
      return None
    skel["vertices"] = skel["vertices"] + minpt.astype(np.float32)
    skel["vertices"] = np.multiply(skel["vertices"], anisotropy, dtype=np.float32)
    return skel

  todo = [s for s in cc_segids if only_cc is None or s in only_cc]
  skeletons = defaultdict(list)
  if parallel is not None and parallel > 1 and len(todo) > 1:
    # the reference's parallel mode: a process pool over connected components (intake.py:344-432);
    # fork keeps cc_labels / all_dbf shared copy-on-write instead of the reference's POSIX shm
    import multiprocessing as mp
    global _POOL_FN
    _POOL_FN = trace_one
    order = sorted(todo, key=lambda s: -counts[s])
    with mp.get_context("fork").Pool(parallel) as pool:
      for segid, skel in zip(order, pool.map(_pool_call, order, chunksize=1)):
        if skel is not None:
          skeletons[remapping[segid]].append(skel)
  else:
    for segid in todo:
      skel = trace_one(segid)
      if skel is not None:
        skeletons[remapping[segid]].append(skel)
  out = {}
  for segid, skels in skeletons.items():
    out[segid] = skel_consolidate(skel_simple_merge(skels))
  if timings is not None:
    timings["preamble_s"] = t1 - t0
    timings["trace_s"] = time.time() - t1
  return out
