"""
Host-side restatement of the reference's fix_borders target selection:

  kimimaro/intake.py:544-585                    compute_border_targets (six faces)
  ext/skeletontricks/skeletontricks.pyx:591-648 find_border_targets
  ext/skeletontricks/skeletontricks.pyx:528-588 compute_centroids
  ext/skeletontricks/skeletontricks.pyx:650-760 compute_tiebreaker_maxima, edgeness, cornerness, distsq

The 2-D connected components and the 2-D EDT of each face run on the device (b2t_ccl26_roots with
sz=1, b2t_edt with ndim=2); what is left is O(face) bookkeeping with the reference's float32 /
float64 expression order, done here in numpy on the six faces (6 * 512^2 voxels for config 3).
Python `set` / `list(set)` containers are kept so that the order in which a label's border targets
are handed to trace() -- and therefore which one becomes the root (intake.py:484-486) -- is inherited
from CPython exactly like the reference inherits it.
"""
import numpy as np

f32 = np.float32


def _distsq(p1x, p1y, p2x, p2y, wx, wy):
  a = f32(wx * f32(p1x - p2x))
  b = f32(wy * f32(p1y - p2y))
  return f32(f32(a * a) + f32(b * b))


def _cornerness(x, y, sx, sy, wx, wy):
  h = f32(0.5)
  # fourth corner uses sx for its y coordinate, as the reference does (pyx:747)
  return min(_distsq(x, y, -h, -h, wx, wy), _distsq(x, y, f32(sx - h), -h, wx, wy),
             _distsq(x, y, f32(sx - h), f32(sy - h), wx, wy), _distsq(x, y, -h, f32(sx - h), wx, wy))


def _edgeness(x, y, sx, sy, wx, wy):
  # evaluated in double (the literal 0.5 is a C double, pyx:725-730), rounded to float on return
  x, y, sx, sy, wx, wy = (float(v) for v in (x, y, sx, sy, wx, wy))
  return f32(min(wx * (x - 0.5), wx * (sx - 0.5 - x), wy * (y - 0.5), wy * (sy - 0.5 - y)))


def tiebreak(px, py, x, y, centx, centy, sx, sy, wx, wy):
  """compute_tiebreaker_maxima (pyx:650-715): closest to the label centroid, then to the plane
  centre, then to a corner, then to an edge, else the previous maximum."""
  px, py, x, y, centx, centy, sx, sy, wx, wy = (f32(v) for v in (px, py, x, y, centx, centy, sx, sy, wx, wy))
  cx = f32(f32(wx * sx) / f32(2.0))
  cy = f32(f32(wy * sy) / f32(2.0))
  for fn in (lambda a, b: _distsq(a, b, centx, centy, wx, wy),
             lambda a, b: _distsq(a, b, cx, cy, wx, wy),
             lambda a, b: _cornerness(a, b, sx, sy, wx, wy),
             lambda a, b: _edgeness(a, b, sx, sy, wx, wy)):
    d1, d2 = fn(px, py), fn(x, y)
    if d2 < d1:
      return (x, y)
    if d1 != d2:
      return (px, py)
  return (px, py)


class Centroids:
  """compute_centroids (pyx:528-588), evaluated lazily per label: float32 running sums accumulated in
  the reference's scan order (x outer, y inner).  One stable sort groups the face by label once."""
  def __init__(self, cc_plane, wx, wy):
    self.wx, self.wy = f32(wx), f32(wy)
    self.sx, self.sy = cc_plane.shape
    flat = np.ascontiguousarray(cc_plane).reshape(-1)        # C order == x outer, y inner
    nz = np.flatnonzero(flat)
    order = np.argsort(flat[nz], kind="stable")              # keeps scan order inside each label
    self.idx = nz[order]
    self.labs = flat[self.idx]
    self.cache = {}

  def __call__(self, label):
    if label in self.cache:
      return self.cache[label]
    a = np.searchsorted(self.labs, label, side="left")
    b = np.searchsorted(self.labs, label, side="right")
    lin = self.idx[a:b]
    xs = (lin // self.sy).astype(np.float32)
    ys = (lin % self.sy).astype(np.float32)
    wx, wy = self.wx, self.wy
    xsum = np.add.accumulate(xs, dtype=np.float32)[-1]
    ysum = np.add.accumulate(ys, dtype=np.float32)[-1]
    ct = f32(lin.size)
    cx = f32(f32(wx * f32(self.sx)) / f32(2))
    cy = f32(f32(wy * f32(self.sy)) / f32(2))
    px = f32(f32(wx * xsum) / ct)
    py = f32(f32(wy * ysum) / ct)
    if not (f32(px - cx) >= 0):
      px = f32(px + wx)
    if not (f32(py - cy) >= 0):
      py = f32(py + wy)
    r = (int(f32(px / wx)), int(f32(py / wy)))
    self.cache[label] = r
    return r


def find_border_targets(dt, cc_plane, wx, wy):
  """find_border_targets (pyx:591-648): {plane label: (x, y) of its DT maximum}, keys in the order the
  raster scan (y outer, x inner) first meets each label on a voxel with non-zero DT."""
  sx, sy = dt.shape
  fdt = np.asarray(dt).reshape(-1, order="F")
  fcc = np.asarray(cc_plane).reshape(-1, order="F")
  sel = np.flatnonzero((fcc != 0) & (fdt != 0))
  pts = {}
  if sel.size == 0:
    return pts
  labs = fcc[sel].astype(np.int64)
  vals = fdt[sel]
  # group by label with ONE stable sort (raster order survives inside each group)
  order = np.argsort(labs, kind="stable")
  sl, sv, si = labs[order], vals[order], sel[order]
  gstart = np.concatenate(([0], np.flatnonzero(np.diff(sl)) + 1))
  gmax = np.maximum.reduceat(sv, gstart)
  glen = np.diff(np.concatenate((gstart, [sl.size])))
  is_max = sv == np.repeat(gmax, glen)
  cand_idx = si[is_max]                        # grouped by label, raster order inside a group
  cand_lab = sl[is_max]
  first_pos = si[gstart]                       # first raster position of every label
  for l in sl[gstart][np.argsort(first_pos, kind="stable")].tolist():  # dict insertion order = first encounter (B.6)
    pts[l] = None
  # Labels with a single maximum are done.  For the others the reference folds compute_tiebreaker_maxima
  # over the candidates in raster order; that comparator is a lexicographic strict order on
  # (distance to the label centroid, to the plane centre, cornerness, edgeness) that keeps the earlier
  # candidate on a full tie, so the fold equals "lexicographic minimum, earliest on ties" -- evaluated here
  # for all candidates of all labels at once with the same float32 / float64 expression order.
  cl, ci = cand_lab, cand_idx
  bounds = np.flatnonzero(np.diff(cl)) + 1
  starts = np.concatenate(([0], bounds))
  ends = np.concatenate((bounds, [cl.size]))
  single = (ends - starts) == 1
  for l, i in zip(cl[starts[single]].tolist(), ci[starts[single]].tolist()):
    pts[l] = (i % sx, i // sx)
  if (~single).any():
    cents = Centroids(cc_plane, wx, wy)
    multi = np.repeat(~single, ends - starts)
    ml, mi = cl[multi], ci[multi]
    px = (mi % sx).astype(np.float32)
    py = (mi // sx).astype(np.float32)
    labs_multi = cl[starts[~single]]
    cxy = np.array([cents(int(l)) for l in labs_multi.tolist()], dtype=np.float32).reshape(-1, 2)
    rep = (ends - starts)[~single]
    centx = np.repeat(cxy[:, 0], rep)
    centy = np.repeat(cxy[:, 1], rep)
    wxf, wyf, fsx, fsy = f32(wx), f32(wy), f32(sx), f32(sy)
    half = f32(0.5)

    def dsq(ax, ay, bx, by):                       # distsq (pyx:750-760), float32 throughout
      u = wxf * (ax - bx)
      v = wyf * (ay - by)
      return u * u + v * v

    c1 = dsq(px, py, centx, centy)
    c2 = dsq(px, py, f32(f32(wxf * fsx) / f32(2.0)), f32(f32(wyf * fsy) / f32(2.0)))
    c3 = np.minimum(np.minimum(dsq(px, py, -half, -half), dsq(px, py, f32(fsx - half), -half)),
                    np.minimum(dsq(px, py, f32(fsx - half), f32(fsy - half)), dsq(px, py, -half, f32(fsx - half))))
    pxd, pyd = px.astype(np.float64), py.astype(np.float64)
    wxd, wyd = float(wxf), float(wyf)
    c4 = np.minimum(np.minimum(wxd * (pxd - 0.5), wxd * (float(sx) - 0.5 - pxd)),
                    np.minimum(wyd * (pyd - 0.5), wyd * (float(sy) - 0.5 - pyd))).astype(np.float32)
    order = np.lexsort((np.arange(ml.size), c4, c3, c2, c1, ml))
    ml_sorted = ml[order]
    firsts = np.concatenate(([0], np.flatnonzero(np.diff(ml_sorted)) + 1))
    for l, i in zip(ml_sorted[firsts].tolist(), mi[order][firsts].tolist()):
      pts[l] = (i % sx, i // sx)
  return pts


def centroid_from_sums(xsum, ysum, count, sx, sy, wx, wy):
  """The tail of compute_centroids (pyx:566-586) from the sequential float32 sums of a label."""
  wx, wy = f32(wx), f32(wy)
  ct = f32(count)
  cx = f32(f32(wx * f32(sx)) / f32(2))
  cy = f32(f32(wy * f32(sy)) / f32(2))
  px = f32(f32(wx * f32(xsum)) / ct)
  py = f32(f32(wy * f32(ysum)) / ct)
  if not (f32(px - cx) >= 0):
    px = f32(px + wx)
  if not (f32(py - cy) >= 0):
    py = f32(py + wy)
  return (int(f32(px / wx)), int(f32(py / wy)))


def centroids_from_sums(xsum, ysum, count, sx, sy, wx, wy):
  """centroid_from_sums for arrays of labels at once -> int64 [n, 2].  Every step is one IEEE float32 operation per
  element (multiply, divide, subtract, add, compare, truncate), which numpy evaluates identically for scalars and arrays."""
  wx, wy = f32(wx), f32(wy)
  ct = np.asarray(count).astype(np.float32)
  cx = f32(f32(wx * f32(sx)) / f32(2))
  cy = f32(f32(wy * f32(sy)) / f32(2))
  px = (wx * np.asarray(xsum, dtype=np.float32)) / ct
  py = (wy * np.asarray(ysum, dtype=np.float32)) / ct
  px = np.where((px - cx) >= 0, px, px + wx)
  py = np.where((py - cy) >= 0, py, py + wy)
  return np.stack([(px / wx).astype(np.int64), (py / wy).astype(np.int64)], axis=1)


def targets_from_candidates(cand_idx, cand_lab, labels, first_pos, xsum, ysum, count, sx, sy, wx, wy):
  """
  find_border_targets (pyx:591-648) when the per-label reductions were done on the device:
    cand_idx / cand_lab  flat (x + sx*y) positions of every voxel that attains its label's DT maximum, in
                         raster order, and their labels
    labels, first_pos    the labels present and the raster position where each is first met
    xsum, ysum, count    per entry of `labels`: sequential float32 coordinate sums (scan order x outer,
                         y inner) and voxel count, for compute_centroids
  Returns {label: (x, y)} with keys in first-encounter order.
  """
  pts = {}
  if len(labels) == 0:
    return pts
  labels = np.asarray(labels)
  for l in labels[np.argsort(np.asarray(first_pos), kind="stable")].tolist():
    pts[l] = None
  slot = {int(l): k for k, l in enumerate(labels.tolist())}
  order = np.argsort(cand_lab, kind="stable")
  cl, ci = np.asarray(cand_lab)[order], np.asarray(cand_idx)[order]
  bounds = np.flatnonzero(np.diff(cl)) + 1
  starts = np.concatenate(([0], bounds))
  ends = np.concatenate((bounds, [cl.size]))
  single = (ends - starts) == 1
  for l, i in zip(cl[starts[single]].tolist(), ci[starts[single]].tolist()):
    pts[l] = (i % sx, i // sx)
  if (~single).any():
    multi = np.repeat(~single, ends - starts)
    ml, mi = cl[multi], ci[multi]
    px = (mi % sx).astype(np.float32)
    py = (mi // sx).astype(np.float32)
    labs_multi = cl[starts[~single]]
    sl = np.array([slot[l] for l in labs_multi.tolist()], dtype=np.int64)
    cxy = centroids_from_sums(np.asarray(xsum)[sl], np.asarray(ysum)[sl], np.asarray(count)[sl], sx, sy, wx,
                              wy).astype(np.float32).reshape(-1, 2)
    rep = (ends - starts)[~single]
    c1, c2, c3, c4 = _criteria(px, py, np.repeat(cxy[:, 0], rep), np.repeat(cxy[:, 1], rep), sx, sy, wx, wy)
    o2 = np.lexsort((np.arange(ml.size), c4, c3, c2, c1, ml))
    ml_sorted = ml[o2]
    firsts = np.concatenate(([0], np.flatnonzero(np.diff(ml_sorted)) + 1))
    for l, i in zip(ml_sorted[firsts].tolist(), mi[o2][firsts].tolist()):
      pts[l] = (i % sx, i // sx)
  return pts


def _criteria(px, py, centx, centy, sx, sy, wx, wy):
  """The four tie-break criteria of compute_tiebreaker_maxima (pyx:650-760) for arrays of candidates, with the
  reference's float32 (distsq, cornerness) and float64 (edgeness) expression order."""
  wxf, wyf, fsx, fsy = f32(wx), f32(wy), f32(sx), f32(sy)
  half = f32(0.5)

  def dsq(ax, ay, bx, by):
    u = wxf * (ax - bx)
    v = wyf * (ay - by)
    return u * u + v * v

  c1 = dsq(px, py, centx, centy)
  c2 = dsq(px, py, f32(f32(wxf * fsx) / f32(2.0)), f32(f32(wyf * fsy) / f32(2.0)))
  c3 = np.minimum(np.minimum(dsq(px, py, -half, -half), dsq(px, py, f32(fsx - half), -half)),
                  np.minimum(dsq(px, py, f32(fsx - half), f32(fsy - half)), dsq(px, py, -half, f32(fsx - half))))
  pxd, pyd = px.astype(np.float64), py.astype(np.float64)
  wxd, wyd = float(wxf), float(wyf)
  c4 = np.minimum(np.minimum(wxd * (pxd - 0.5), wxd * (float(sx) - 0.5 - pxd)),
                  np.minimum(wyd * (pyd - 0.5), wyd * (float(sy) - 0.5 - pyd))).astype(np.float32)
  return c1, c2, c3, c4


def plane_mapping(plane, cc_plane):
  """get_mapping (pyx:490-525) on a face: {plane component id: volume cc label}.  (All-host variant, kept for
  the CPU tests; the engine builds this table on the device.)"""
  fcc = np.asarray(cc_plane).reshape(-1)
  fpl = np.asarray(plane).reshape(-1)
  _, first = np.unique(fcc, return_index=True)
  return {int(fcc[i]): int(fpl[i]) for i in first}
