"""
Host side of the B200 TEASAR engine: the conductor that kimimaro/intake.py:58-221,434-593 and
kimimaro/trace.py:36-194 are in the reference, re-hosted so that every per-voxel step is one
launch over the WHOLE volume / ALL labels (field.cu) and the sequential path loop is one
device-resident launch (trace.cu).  torch tensors are device buffers only.

Data layout in HBM (per voxel of the [sx,sy,sz] Fortran-ordered volume, V voxels):
  labels  L bytes   input ids                      cc      4 B  connected-component id (0 = background)
  dbf     4 B  distance-to-boundary (K1)           dist    4 B  DAF, then railroad scratch (+inf at rest)
  pdrf    4 B  penalised distance field            claim   8 B  ~0 = valid; 0 = invalidated; else pending (dist,seed)
  stamp   4 B  frontier de-duplication flags
plus per foreground voxel: keys 8 B (DAF-bucketed target list), 88 B of queue scratch, path pool.
"""
import ctypes
import os
import time
from collections import defaultdict

import numpy as np
import torch

from . import _lib, border
from ._lib import B2TError, c_f32, c_i64, c_int, c_vp, check, lib, stream_ptr
from .ops import edt

c_u32 = ctypes.c_uint32
c_u64 = ctypes.c_uint64

_lib.declare("b2t_label_stats", [c_vp, c_vp, c_i64, c_i64, c_i64, c_u32, c_vp, c_vp, c_vp, c_vp, c_vp])
_lib.declare("b2t_edf_multi", [c_vp, c_i64, c_i64, c_i64, c_f32, c_f32, c_f32, c_vp, c_u32, c_f32, c_u32,
                               c_vp, c_vp, c_vp, c_vp, c_u64, c_vp, c_vp])
_lib.declare("b2t_edf_labels", [c_vp, c_i64, c_i64, c_i64, c_f32, c_f32, c_f32, c_vp, c_u32, c_u32, c_vp, c_vp, c_vp, c_vp,
                                c_vp, c_vp])
_lib.declare("b2t_field_argmax", [c_vp, c_vp, c_i64, c_i64, c_i64, c_u32, c_vp, c_vp])
_lib.declare("b2t_pdrf_and_buckets", [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_u32, c_vp, c_vp,
                                      c_vp, c_u32, c_f32, c_f32, c_int, c_vp, c_vp, c_vp, c_vp])
_lib.declare("b2t_trace_batch", [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_f32, c_f32, c_f32,
                                 c_vp, c_int, c_f32, c_f32, c_f32, c_f32, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                                 c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_f32, c_vp, c_u64, c_u64, c_int, c_vp, c_vp])
_lib.declare("b2t_trace_heap_words", [c_u64, c_u64], c_u64)
_lib.declare("b2t_trace_team_bytes", [], c_u64)
_lib.declare("b2t_trace_scratch_words", [c_u64], c_u64)
_lib.declare("b2t_ccl26_roots", [c_vp, c_int, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp])
_lib.declare("b2t_ccl_relabel", [c_vp, c_vp, c_u64, c_vp])
_lib.declare("b2t_invalidate_ball", [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_f32, c_f32, c_f32, c_vp, c_u32, c_f32, c_f32,
                                     c_vp, c_vp, c_u64, c_vp, c_vp])
_lib.declare("b2t_face_stats", [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp])
_lib.declare("b2t_invalidate_ball_single", [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_f32, c_f32, c_f32, c_u32, c_f32, c_f32,
                                            c_vp, c_vp, c_vp, c_vp, c_vp])
_lib.declare("b2t_segment_seqsum", [c_vp, c_vp, c_vp, c_u32, c_vp, c_vp, c_vp])
_lib.declare("b2t_set_launch_limits", [c_int, c_int])
_lib.declare("b2t_gather_paths", [c_vp, c_vp, c_vp, c_vp, c_u32, c_vp, c_vp, c_vp, c_vp])
_lib.declare("b2t_assemble_group_cap", [], c_u32)
_lib.declare("b2t_assemble", [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_u32, c_vp, c_i64, c_i64, c_i64, c_f32, c_f32, c_f32,
                              c_f32, c_f32, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp])

NBUCKETS = 256
NONE = 0xFFFFFFFF

# trace()'s own defaults (kimimaro/trace.py:38-43), which differ from DEFAULT_TEASAR_PARAMS
TRACE_DEFAULTS = dict(scale=10, const=10, soma_detection_threshold=1100, soma_acceptance_threshold=4000,
                      pdrf_scale=5000, pdrf_exponent=16, soma_invalidation_scale=0.5, soma_invalidation_const=0,
                      max_paths=None)


def _p(t):
  return c_vp(t.data_ptr()) if t is not None else c_vp(0)


def _dev(a):
  t = torch.from_numpy(np.ascontiguousarray(a))
  return t.cuda(non_blocking=True)


# ------------------------------------------------------------------------------------------------
# device-side steps
# ------------------------------------------------------------------------------------------------
def connected_components(d_labels, shape):
  """cc3d.connected_components(labels) (utility.py:77): returns (cc int32 [V], n_cc)."""
  sx, sy, sz = shape
  V = sx * sy * sz
  parent = torch.empty(V, dtype=torch.int32, device=d_labels.device)
  is_root = torch.empty(V, dtype=torch.uint8, device=d_labels.device)
  check(lib().b2t_ccl26_roots(_p(d_labels), c_int(d_labels.element_size()), c_i64(sx), c_i64(sy), c_i64(sz),
                              _p(parent), _p(is_root), stream_ptr()), "b2t_ccl26_roots")
  rank = torch.cumsum(is_root, 0, dtype=torch.int32)       # plumbing: rank of every root in raster order
  n_cc = int(rank[-1].item()) if V > 0 else 0
  del is_root
  check(lib().b2t_ccl_relabel(_p(parent), _p(rank), c_u64(V), stream_ptr()), "b2t_ccl_relabel")
  return parent, n_cc


def _ccl_nosync(d_labels, shape):
  """connected_components without reading the component count back (faces only need the labels)."""
  sx, sy, sz = shape
  V = sx * sy * sz
  parent = torch.empty(V, dtype=torch.int32, device=d_labels.device)
  is_root = torch.empty(V, dtype=torch.uint8, device=d_labels.device)
  check(lib().b2t_ccl26_roots(_p(d_labels), c_int(d_labels.element_size()), c_i64(sx), c_i64(sy), c_i64(sz),
                              _p(parent), _p(is_root), stream_ptr()), "b2t_ccl26_roots")
  rank = torch.cumsum(is_root, 0, dtype=torch.int32)
  check(lib().b2t_ccl_relabel(_p(parent), _p(rank), c_u64(V), stream_ptr()), "b2t_ccl_relabel")
  return parent


def label_stats(d_cc, d_dbf, shape, n):
  sx, sy, sz = shape
  dev = d_cc.device
  count = torch.empty(n + 1, dtype=torch.int32, device=dev)
  bbox = torch.empty((n + 1) * 6, dtype=torch.int32, device=dev)
  dbfmax = torch.empty(n + 1, dtype=torch.float32, device=dev)
  first = torch.empty(n + 1, dtype=torch.int32, device=dev)
  check(lib().b2t_label_stats(_p(d_cc), _p(d_dbf), c_i64(sx), c_i64(sy), c_i64(sz), c_u32(n), _p(count), _p(bbox),
                              _p(dbfmax), _p(first), stream_ptr()), "b2t_label_stats")
  return count, bbox, dbfmax, first


class Workspace:
  """Per-volume work fields (allocated once per skeletonize call, reused by both field sweeps)."""
  def __init__(self, V, n_fg, dev):
    self.V, self.n_fg = V, n_fg
    self.dist = torch.empty(V, dtype=torch.float32, device=dev)
    self.stamp = torch.empty(V, dtype=torch.int32, device=dev)
    self.queue = torch.empty(2 * max(n_fg, 1), dtype=torch.int32, device=dev)
    self.ctrl = torch.zeros(16, dtype=torch.int32, device=dev)


def edf_multi(d_cc, shape, anisotropy, d_sources, n_sources, ws, free_space=None, node_weights=None):
  """dijkstra3d.euclidean_distance_field for all participating labels at once -> ws.dist.
  node_weights (a [V] float tensor) turns it into dijkstra3d.parental_field(node_weights, source)."""
  sx, sy, sz = shape
  ws.dist.fill_(float("inf"))
  ws.stamp.zero_()
  fsr, fss = (0.0, 0) if free_space is None else (float(free_space[0]), int(free_space[1]))
  check(lib().b2t_edf_multi(_p(d_cc), c_i64(sx), c_i64(sy), c_i64(sz), c_f32(anisotropy[0]), c_f32(anisotropy[1]),
                            c_f32(anisotropy[2]), _p(d_sources), c_u32(n_sources), c_f32(fsr), c_u32(fss),
                            _p(node_weights), _p(ws.dist), _p(ws.stamp), _p(ws.queue), c_u64(ws.queue.numel() // 2), _p(ws.ctrl),
                            stream_ptr()), "b2t_edf_multi")


# labels above EDF_TEAM_MIN voxels get a thread-block cluster in b2t_edf_labels (at most EDF_TEAM_MAX of them), labels above
# EDF_GRID_MIN the grid-wide cooperative sweep (b2t_edf_multi): one label that size keeps every SM busy by itself
EDF_TEAM_MIN = int(os.environ.get("B2T_EDF_TEAM_MIN", 32768))
EDF_TEAM_MAX = int(os.environ.get("B2T_EDF_TEAM_MAX", 96))
EDF_GRID_MIN = int(os.environ.get("B2T_EDF_GRID_MIN", 4 << 20))


def edf_labels(d_cc, shape, anisotropy, sources, segids, n_fg, ws, node_weights=None):
  """dijkstra3d.euclidean_distance_field (or parental_field with node_weights) for many labels, each label with its own
  relaxation rounds (b2t_edf_labels) -> ws.dist.  sources / segids / n_fg: one entry per label."""
  sx, sy, sz = shape
  sources = np.asarray(sources, dtype=np.int64)
  n_fg = np.asarray(n_fg, dtype=np.int64)
  segids = np.asarray(segids, dtype=np.int64)
  ws.dist.fill_(float("inf"))
  ws.stamp.zero_()
  giant = n_fg > EDF_GRID_MIN
  if giant.any() or os.environ.get("B2T_EDF_LABELS", "1") == "0":
    sel = np.flatnonzero(giant) if os.environ.get("B2T_EDF_LABELS", "1") != "0" else np.arange(n_fg.size)
    src = _dev(sources[sel].astype(np.uint32).view(np.int32))
    check(lib().b2t_edf_multi(_p(d_cc), c_i64(sx), c_i64(sy), c_i64(sz), c_f32(anisotropy[0]), c_f32(anisotropy[1]),
                              c_f32(anisotropy[2]), _p(src), c_u32(int(sel.size)), c_f32(0.0), c_u32(0),
                              _p(node_weights), _p(ws.dist), _p(ws.stamp), _p(ws.queue), c_u64(ws.queue.numel() // 2),
                              _p(ws.ctrl), stream_ptr()), "b2t_edf_multi")
    rest = np.setdiff1d(np.arange(n_fg.size), sel)
    if rest.size == 0:
      return
    if giant.any():
      ws.stamp.zero_()             # round numbers restart per label; the giant labels' stamps are done with
    sources, segids, n_fg = sources[rest], segids[rest], n_fg[rest]
  order = np.argsort(-n_fg, kind="stable")
  n = int(order.size)
  tab = np.zeros((n, 4), dtype=np.uint32)
  tab[:, 0] = sources[order]
  tab[:, 1] = segids[order]
  tab[:, 2] = n_fg[order]
  tab[:, 3] = np.concatenate(([0], np.cumsum(n_fg[order])[:-1]))
  n_team = min(int((n_fg > EDF_TEAM_MIN).sum()), EDF_TEAM_MAX)
  if 2 * int(n_fg.sum()) > ws.queue.numel():
    raise B2TError("edf_labels: queue too small for these labels")
  d_tab = _dev(tab.view(np.int32))
  ctrl = torch.zeros(4 * max(n_team, 1), dtype=torch.int32, device=d_cc.device)
  check(lib().b2t_edf_labels(_p(d_cc), c_i64(sx), c_i64(sy), c_i64(sz), c_f32(anisotropy[0]), c_f32(anisotropy[1]),
                             c_f32(anisotropy[2]), _p(d_tab), c_u32(n), c_u32(n_team), _p(node_weights), _p(ws.dist),
                             _p(ws.stamp), _p(ws.queue), _p(ctrl), stream_ptr()), "b2t_edf_labels")
  ws.keep = (d_tab, ctrl)          # alive until the stream gets there


def field_argmax(d_cc, d_dist, shape, n):
  """per label (max finite distance, smallest index attaining it) -> (values f32[n+1], index i64[n+1])."""
  return field_argmax_read(field_argmax_launch(d_cc, d_dist, shape, n))


def field_argmax_launch(d_cc, d_dist, shape, n):
  sx, sy, sz = shape
  best = torch.empty(n + 1, dtype=torch.int64, device=d_cc.device)
  check(lib().b2t_field_argmax(_p(d_cc), _p(d_dist), c_i64(sx), c_i64(sy), c_i64(sz), c_u32(n), _p(best),
                               stream_ptr()), "b2t_field_argmax")
  return best


def field_argmax_read(best):
  b = best.cpu().numpy().view(np.uint64)
  vals = (b >> np.uint64(32)).astype(np.uint32).view(np.float32)
  idx = (np.uint64(0xFFFFFFFF) - (b & np.uint64(0xFFFFFFFF))).astype(np.int64)
  idx[b == 0] = -1
  return vals, idx


class RootSweep:
  """find_root (trace.py:128-129, 291-308) for the labels of the main arena, started BEFORE the border targets are
  computed and collected after: the sweep is a latency-bound chain of per-label rounds that leaves most of the GPU idle,
  compute_border_targets is a sequence of small launches and host reads -- on two streams they overlap (5.7 + 5.9 ms in
  a row on synthetic-512).  Labels that get their root from a border target simply ignore the result."""
  def __init__(self, d_cc, shape, anisotropy, first, segids, n_fg, n_rows):
    V = shape[0] * shape[1] * shape[2]
    self.ws = Workspace(V, int(np.asarray(n_fg).sum()), d_cc.device)
    self.segids = np.asarray(segids, dtype=np.int64)
    self.side = None
    if d_cc.device.type == "cuda":
      self.main = torch.cuda.current_stream()
      self.side = torch.cuda.Stream()
      self.side.wait_stream(self.main)
      with torch.cuda.stream(self.side):
        edf_labels(d_cc, shape, anisotropy, first, segids, n_fg, self.ws)
        self.best = field_argmax_launch(d_cc, self.ws.dist, shape, n_rows)
    else:
      edf_labels(d_cc, shape, anisotropy, first, segids, n_fg, self.ws)
      self.best = field_argmax_launch(d_cc, self.ws.dist, shape, n_rows)

  def roots(self):
    """linear index of the root of every label handed to the constructor (same order)"""
    if self.side is not None:
      self.main.wait_stream(self.side)         # the workspace goes back to the main stream (the DAF sweep reuses it)
      self.side.synchronize()
    _, idx = field_argmax_read(self.best)
    return idx[self.segids]


# ------------------------------------------------------------------------------------------------
# fix_borders (intake.py:544-585)
# ------------------------------------------------------------------------------------------------
def compute_border_targets(d_cc, shape, anisotropy, shard=None):
  """intake.py:544-585.  shard = (rank, world_size, group) of a multi-GPU run: face f is computed on rank f % world_size
  only, the per-face target lists are exchanged (one all_gather_object of a few hundred tuples) and every rank replays
  the insertions in face order -- the per-label SETS come out with the same insertion history, hence the same iteration
  order (rule B.6), as on one GPU.  The six faces cost 6-9 ms of launches and host work that does not shrink otherwise."""
  sx, sy, sz = shape
  cc3 = d_cc.view(sz, sy, sx)
  faces = (
    (cc3[0, :, :], (sx, sy), (0, 1), lambda x, y: (x, y, 0)),
    (cc3[sz - 1, :, :], (sx, sy), (0, 1), lambda x, y: (x, y, sz - 1)),
    (cc3[:, 0, :], (sx, sz), (0, 2), lambda x, z: (x, 0, z)),
    (cc3[:, sy - 1, :], (sx, sz), (0, 2), lambda x, z: (x, sy - 1, z)),
    (cc3[:, :, 0], (sy, sz), (1, 2), lambda y, z: (0, y, z)),
    (cc3[:, :, sx - 1], (sy, sz), (1, 2), lambda y, z: (sx - 1, y, z)),
  )
  # Per face, on the device and without a host round trip: 2-D connected components, 2-D EDT, then b2t_face_stats (the
  # per-component reductions of find_border_targets / compute_centroids / get_mapping, compacted).  The six faces are
  # queued back to back; the host then reads the two counters of every face at once and the few thousand candidate /
  # record words they announce, and applies the reference's tie-break order (border.targets_from_candidates).
  L = lib()
  dev = d_cc.device
  pmax = max(sx * sy, sx * sz, sy * sz)
  tab = torch.empty(6 * (pmax + 1), dtype=torch.int32, device=dev)
  cand = torch.empty((6, 2 * pmax), dtype=torch.int32, device=dev)
  rec = torch.empty((6, 7 * pmax), dtype=torch.int32, device=dev)
  counts = torch.zeros((6, 2), dtype=torch.int32, device=dev)
  planes = []
  mine = [f for f in range(6) if shard is None or f % shard[1] == shard[0]]
  for f, (face, pshape, dims, rotatefn) in enumerate(faces):
    if f not in mine:
      planes.append(None)
      continue
    wx, wy = float(anisotropy[dims[0]]), float(anisotropy[dims[1]])
    p0, p1 = int(pshape[0]), int(pshape[1])
    plane = face.contiguous().view(-1)                     # flat Fortran order of the 2-D plane: x + p0*y
    cc_plane = _ccl_nosync(plane, (p0, p1, 1))             # 8-connected in 2-D (intake.py:564)
    dt = edt(cc_plane, pshape, anisotropy=(wx, wy), black_border=True)       # intake.py:565
    check(L.b2t_face_stats(_p(cc_plane), _p(dt), _p(plane), c_i64(p0), c_i64(p1), _p(tab), _p(cand[f]), _p(rec[f]),
                           _p(counts[f]), stream_ptr()), "b2t_face_stats")
    planes.append((cc_plane, p0, p1, wx, wy, rotatefn))
  h_counts = counts.cpu().numpy()                          # the one synchronising read
  parts = []
  for f in mine:
    parts.append(cand[f, :2 * int(h_counts[f, 0])])
    parts.append(rec[f, :7 * int(h_counts[f, 1])])
  h_all = torch.cat(parts).cpu().numpy().view(np.uint32) if parts else np.zeros(0, np.uint32)
  face_lists = {f: [] for f in mine}                       # per face: (volume label, point) in insertion order
  o = 0
  for f in mine:
    cc_plane, p0, p1, wx, wy, rotatefn = planes[f]
    nc, nr = int(h_counts[f, 0]), int(h_counts[f, 1])
    c = h_all[o:o + 2 * nc].reshape(-1, 2).astype(np.int64)
    o += 2 * nc
    r = h_all[o:o + 7 * nr].reshape(-1, 7).astype(np.int64)
    o += 7 * nr
    if nr == 0:
      continue
    order = np.lexsort((c[:, 0], c[:, 1]))                 # by component, raster order inside (find_border_targets' scan)
    cand_idx, cand_lab = c[order, 0], c[order, 1]
    labels, first_pos, cnt, vol_of = r[:, 0], r[:, 2], r[:, 3], r[:, 6]
    sumx, sumy = r[:, 4].astype(np.float32), r[:, 5].astype(np.float32)
    # integer sums equal the reference's sequential float32 sums while every partial sum is exact in float32; a larger
    # component whose maximum is tied (only those use the centroid) is summed again in the reference's order
    big = cnt * max(p0 - 1, p1 - 1, 1) >= (1 << 24)
    if big.any():
      tied = set(np.unique(cand_lab[np.flatnonzero(np.diff(cand_lab) == 0)]).tolist())
      redo = [k for k in np.flatnonzero(big).tolist() if int(labels[k]) in tied]
      if redo:
        h_plane = cc_plane.cpu().numpy().view(np.uint32).reshape(p1, p0).T      # [x, y]
        for k in redo:
          xs, ys = np.nonzero(h_plane == labels[k])        # C order of [x, y]: x outer, y inner (pyx:551-560)
          sumx[k] = np.add.accumulate(xs.astype(np.float32), dtype=np.float32)[-1]
          sumy[k] = np.add.accumulate(ys.astype(np.float32), dtype=np.float32)[-1]
    plane_targets = border.targets_from_candidates(cand_idx, cand_lab, labels, first_pos, sumx, sumy, cnt,
                                                   p0, p1, wx, wy)
    vol = {int(l): int(v) for l, v in zip(labels.tolist(), vol_of.tolist())}
    for label, pt in plane_targets.items():
      face_lists[f].append((vol[label], rotatefn(int(pt[0]), int(pt[1]))))
  if shard is not None and shard[1] > 1:
    import torch.distributed as dist
    gathered = [None] * shard[1]
    dist.all_gather_object(gathered, face_lists, group=shard[2])
    face_lists = {}
    for part in gathered:
      face_lists.update(part)
  target_list = defaultdict(set)
  for f in range(6):
    for label, pt in face_lists.get(f, []):
      target_list[label].add(tuple(pt))
  out = {}
  for label, pts in target_list.items():
    out[label] = np.array(list(pts), dtype=np.uint32)
  return out


# ------------------------------------------------------------------------------------------------
# the trace of all labels of one arena
# ------------------------------------------------------------------------------------------------
DESC_DTYPE = np.dtype([
  ("segid", "<u4"), ("root", "<u4"), ("n_fg", "<u4"), ("region_off", "<u4"), ("path_off", "<u4"),
  ("path_cap", "<u4"), ("tb_off", "<u4"), ("tb_n", "<u4"), ("ta_off", "<u4"), ("ta_n", "<u4"),
  ("max_paths", "<u4"), ("soma_mode", "<u4"), ("soma_radius", "<f4"), ("bucket_row", "<u4"),
  ("soma_done", "<u4"), ("pre_invalid", "<u4"), ("bbox_x0", "<u4"), ("bbox_x1", "<u4"), ("single_path", "<u4"),
  ("reserved1", "<u4"),
])
assert DESC_DTYPE.itemsize == 80


class Jobs:
  """The per-label arguments of trace() (kimimaro/intake.py:494-504) for every label of one arena, as arrays.
  root == -1 means "find a root" (trace.py:128-129); tb / ta map a job index to its list of manual targets
  (linear voxel indices) and only hold the labels that have any."""
  def __init__(self, segid, n_fg, first, root, dbf_max, tb=None, ta=None, soma_mode=None, soma_radius=None,
               free_space=None, bbox_x=None, daf_source=None):
    self.segid = np.asarray(segid, dtype=np.int64)
    n = self.segid.size
    self.n_fg = np.asarray(n_fg, dtype=np.int64)
    self.first = np.asarray(first, dtype=np.int64)
    self.root = np.asarray(root, dtype=np.int64).copy()
    self.dbf_max = np.asarray(dbf_max, dtype=np.float32)
    self.tb = tb if tb is not None else {}
    self.ta = ta if ta is not None else {}
    self.soma_mode = np.zeros(n, dtype=bool) if soma_mode is None else np.asarray(soma_mode, dtype=bool)
    self.soma_radius = np.zeros(n, dtype=np.float32) if soma_radius is None else np.asarray(soma_radius, dtype=np.float32)
    self.free_space = np.zeros(n, dtype=np.float32) if free_space is None else np.asarray(free_space, dtype=np.float32)
    # x extent (inclusive) of every label's bounding box: the array the reference's invalidation runs on
    self.bbox_x = None if bbox_x is None else np.asarray(bbox_x, dtype=np.int64).reshape(n, 2)
    # source of the DAF when it is not the root (point_to_point, trace.py:358-390: DAF from `start`, path field from `end`)
    self.daf_source = None if daf_source is None else np.asarray(daf_source, dtype=np.int64)
    self.single_path = False    # True: every job yields exactly one path, to its first manual target, nothing is invalidated

  def __len__(self):
    return int(self.segid.size)


def compute_M_array(dbf_max):
  """compute_M for every label.  Deliberately a loop over numpy SCALARS: the reference evaluates
  `f(1 / (dbf_max ** 1.01))` on a np.float32 scalar (trace.py:336), and numpy's vectorised float32 power
  (SIMD) rounds differently from the scalar path in ~17 % of the cases -- one ulp of M is enough to move a path."""
  with np.errstate(all="ignore"):
    return np.array([np.float32(1 / (d ** 1.01)) for d in np.asarray(dbf_max, dtype=np.float32)], dtype=np.float32)


def trace_arena_start(d_cc, d_dbf, shape, anisotropy, jobs, params, n_rows, timings=None, pool_scale=1, ws=None):
  """
  Everything up to and including the (asynchronous) launch of the path-loop kernel; returns the state that
  trace_arena_finish() needs.  jobs: a Jobs table.  n_rows: rows of the (label x bucket) tables minus one (= max cc id in this arena).
  """
  sx, sy, sz = shape
  V = sx * sy * sz
  dev = d_cc.device
  n_jobs = len(jobs)
  L = lib()
  tmark = time.perf_counter()

  def lap(name):
    nonlocal tmark
    if timings is not None:
      torch.cuda.synchronize()
      now = time.perf_counter()
      timings[name] = timings.get(name, 0.0) + (now - tmark)
      tmark = now

  n_fg_total = int(jobs.n_fg.sum())
  if ws is None or ws.V != V or ws.n_fg < n_fg_total:
    ws = Workspace(V, n_fg_total, dev)

  # ---- roots (trace.py:128-129, 291-308): one field sweep for every label that has no root yet ----
  need = np.flatnonzero(jobs.root < 0)
  if need.size:
    edf_labels(d_cc, shape, anisotropy, jobs.first[need], jobs.segid[need], jobs.n_fg[need], ws)
    _, idx = field_argmax(d_cc, ws.dist, shape, n_rows)
    jobs.root[need] = idx[jobs.segid[need]]
  lap("find_root")

  # ---- DAF from the roots, all labels at once (trace.py:139-145) ----
  has_free = jobs.free_space > 0
  if has_free.any() and n_jobs != 1:
    raise B2TError("free-space seeding (soma mode) needs an arena of its own")
  src = _dev(jobs.root.astype(np.uint32).view(np.int32))
  if has_free.any():
    edf_multi(d_cc, shape, anisotropy, src, 1, ws, free_space=(float(jobs.free_space[0]), int(jobs.root[0])))
  else:
    edf_labels(d_cc, shape, anisotropy, jobs.root if jobs.daf_source is None else jobs.daf_source, jobs.segid, jobs.n_fg, ws)
  daf_best = field_argmax_launch(d_cc, ws.dist, shape, n_rows)
  # host work that needs nothing from the sweep runs while it is in flight: M per label (a loop over numpy scalars, see
  # compute_M_array), the tables and the buffers of the next step
  M = np.zeros(n_rows + 1, dtype=np.float32)
  inv = np.zeros(n_rows + 1, dtype=np.float32)
  row = np.full(n_rows + 1, NONE, dtype=np.uint32)          # cc id -> row of the (label x bucket) tables: one row per job
  M[jobs.segid] = compute_M_array(jobs.dbf_max)
  row[jobs.segid] = np.arange(n_jobs, dtype=np.uint32)
  d_M, d_row = _dev(M), _dev(row.view(np.int32))
  ntab = max(n_jobs, 1) * NBUCKETS
  hist = torch.empty(ntab + 1, dtype=torch.int32, device=dev)
  cursor = torch.empty(ntab + 1, dtype=torch.int32, device=dev)
  keys = torch.empty(max(n_fg_total, 1), dtype=torch.int64, device=dev)
  pdrf = torch.empty(V, dtype=torch.float32, device=dev)
  claim = torch.empty(V, dtype=torch.int64, device=dev)
  maxdaf, target_idx = field_argmax_read(daf_best)
  lap("daf")

  # ---- PDRF + work fields + target buckets (trace.py:147-148, 315-356; pyx:995-1006) ----
  md = maxdaf[jobs.segid].astype(np.float32)
  with np.errstate(all="ignore"):
    inv[jobs.segid] = np.where(md != 0, np.float32(1) / md, np.float32(0)).astype(np.float32)   # trace.py:352-354
  d_inv = _dev(inv)
  check(L.b2t_pdrf_and_buckets(_p(d_cc), _p(d_dbf), _p(ws.dist), _p(pdrf), _p(claim), c_i64(sx), c_i64(sy),
                               c_i64(sz), c_u32(n_rows), _p(d_M), _p(d_inv), _p(d_row), c_u32(max(n_jobs, 1)),
                               c_f32(params["pdrf_scale"]), c_f32(params["pdrf_exponent"]), c_int(NBUCKETS),
                               _p(hist), _p(cursor), _p(keys), stream_ptr()), "b2t_pdrf_and_buckets")
  lap("pdrf")
  fix_branching = bool(params.get("fix_branching", True))
  if not fix_branching:
    # parents = dijkstra3d.parental_field(PDRF, root) (trace.py:154-155), every label in one sweep; the
    # distances stay in ws.dist and the path kernel derives parents from them (rule T3)
    edf_labels(d_cc, shape, anisotropy, jobs.root, jobs.segid, jobs.n_fg, ws, node_weights=pdrf)
    lap("parental_field")

  # ---- the path loop for every label (trace.py:196-267) ----
  order = np.argsort(-jobs.n_fg, kind="stable")                       # largest labels first
  desc = np.zeros(n_jobs, dtype=DESC_DTYPE)
  desc["segid"] = jobs.segid[order]
  desc["root"] = jobs.root[order]
  desc["n_fg"] = jobs.n_fg[order]
  desc["soma_mode"] = jobs.soma_mode[order]
  desc["soma_radius"] = jobs.soma_radius[order]
  desc["single_path"] = 1 if getattr(jobs, "single_path", False) else 0
  if jobs.bbox_x is not None:
    desc["bbox_x0"], desc["bbox_x1"] = jobs.bbox_x[order, 0], jobs.bbox_x[order, 1]
  else:
    desc["bbox_x0"], desc["bbox_x1"] = 0, sx - 1
  # manual targets: labels without any get the DAF arg-max appended unless in soma mode (trace.py:171-172)
  tb_n = np.zeros(n_jobs, dtype=np.int64)
  ta_n = np.zeros(n_jobs, dtype=np.int64)
  for i, lst in jobs.tb.items():
    tb_n[i] = len(lst)
  for i, lst in jobs.ta.items():
    ta_n[i] = len(lst)
  auto = (tb_n == 0) & ~jobs.soma_mode
  tb_n_eff = tb_n + auto
  cnt = (tb_n_eff + ta_n)[order]
  starts = np.concatenate(([0], np.cumsum(cnt)))
  targets = np.zeros(int(starts[-1]) + 1, dtype=np.int64)
  desc["tb_off"] = starts[:-1]
  desc["tb_n"] = tb_n_eff[order]
  desc["ta_off"] = starts[:-1] + tb_n_eff[order]
  desc["ta_n"] = ta_n[order]
  slot_of = np.empty(n_jobs, dtype=np.int64)
  slot_of[order] = np.arange(n_jobs)
  auto_idx = np.flatnonzero(auto)
  targets[starts[slot_of[auto_idx]]] = target_idx[jobs.segid[auto_idx]]
  for i, lst in jobs.tb.items():
    if len(lst):
      o = starts[slot_of[i]]
      targets[o:o + len(lst)] = lst
  for i, lst in jobs.ta.items():
    if len(lst):
      o = starts[slot_of[i]] + tb_n_eff[i]
      targets[o:o + len(lst)] = lst
  nfg = desc["n_fg"].astype(np.int64)
  # Path pool.  fix_branching=True: every path voxel but the rail end is new to the rail, so sum(len - 1) <= n_fg, and a
  # label has at most n_fg + (manual targets) paths of len + 1 slots each: 3 * n_fg + 2 * cnt bounds the pool.
  # fix_branching=False: path_from_parents writes the whole root -> target walk for every target (about tips x depth);
  # no bound short of n_fg^2 exists, so the pool starts at the same size and labels that report B2T_ERR_CAPACITY are
  # traced again with a larger one (trace_arena_finish).
  caps = (3 * nfg + 2 * cnt + 64) * int(pool_scale)
  desc["region_off"] = np.concatenate(([0], np.cumsum(nfg)[:-1]))
  desc["path_off"] = np.concatenate(([0], np.cumsum(caps)[:-1]))
  desc["path_cap"] = caps
  desc["max_paths"] = NONE if params["max_paths"] is None else int(params["max_paths"])
  desc["bucket_row"] = order                                  # row of the job in the bucket tables (row[segid] above)
  region = int(nfg.sum())
  path_off = int(caps.sum())
  if path_off >= 2 ** 32 or region >= 2 ** 32:
    raise B2TError(f"arena too large for 32-bit pool offsets: path pool {path_off} slots, {region} foreground voxels")
  SCR = 22                                                    # u32 of scratch per voxel (b2t_trace_scratch_words)
  scratch = torch.empty(int(L.b2t_trace_scratch_words(c_u64(max(region, 1)))), dtype=torch.int32, device=dev)
  # soma labels: the one-off ball around the root (trace.py:160-168) is far too large for one CTA
  for slot in np.flatnonzero(desc["soma_mode"]).tolist():
    if desc[slot]["soma_mode"]:
      n = int(desc[slot]["n_fg"])
      base = SCR * int(desc[slot]["region_off"])
      if os.environ.get("B2T_BALL_CCL", "1") == "1":
        # one seed: the claimed set is a connected component, no frontier sweep needed (b2t_invalidate_ball_single)
        mark = torch.empty(V, dtype=torch.uint8, device=dev)
        parent = torch.empty(V, dtype=torch.int32, device=dev)
        is_root = torch.empty(V, dtype=torch.uint8, device=dev)
        check(L.b2t_invalidate_ball_single(_p(d_cc), _p(d_dbf), _p(claim), c_i64(sx), c_i64(sy), c_i64(sz),
                                           c_f32(anisotropy[0]), c_f32(anisotropy[1]), c_f32(anisotropy[2]),
                                           c_u32(int(desc[slot]["root"])), c_f32(params["soma_invalidation_scale"]),
                                           c_f32(params["soma_invalidation_const"]), _p(mark), _p(parent), _p(is_root),
                                           _p(ws.ctrl), stream_ptr()), "b2t_invalidate_ball_single")
      else:
        seeds = _dev(np.array([desc[slot]["root"]], dtype=np.uint32).view(np.int32))
        check(L.b2t_invalidate_ball(_p(d_cc), _p(d_dbf), _p(claim), c_i64(sx), c_i64(sy), c_i64(sz),
                                    c_f32(anisotropy[0]), c_f32(anisotropy[1]), c_f32(anisotropy[2]), _p(seeds), c_u32(1),
                                    c_f32(params["soma_invalidation_scale"]), c_f32(params["soma_invalidation_const"]),
                                    c_vp(scratch.data_ptr() + 4 * base), c_vp(scratch.data_ptr() + 4 * (base + 2 * n)),
                                    c_u64(n), _p(ws.ctrl), stream_ptr()), "b2t_invalidate_ball")
      desc[slot]["soma_done"] = 1
      desc[slot]["pre_invalid"] = int(ws.ctrl[6].item())
  lap("soma_ball")
  d_desc = _dev(desc.view(np.uint8))
  d_targets = _dev(targets.astype(np.uint32).view(np.int32))
  paths = torch.empty(max(path_off, 1), dtype=torch.int32, device=dev)
  out_len = torch.zeros(n_jobs, dtype=torch.int32, device=dev)
  out_np = torch.zeros(n_jobs, dtype=torch.int32, device=dev)
  out_status = torch.zeros(n_jobs, dtype=torch.int32, device=dev)
  out_stats = torch.zeros(4 * n_jobs, dtype=torch.int32, device=dev)
  counter = torch.zeros(1, dtype=torch.int32, device=dev)
  launch = dict(d_cc=d_cc, d_dbf=d_dbf, pdrf=pdrf, ws=ws, claim=claim, shape=shape, anisotropy=anisotropy, params=params,
                fix_branching=fix_branching, keys=keys, hist=hist, cursor=cursor, scratch=scratch, d_targets=d_targets,
                nfg_sorted=nfg)
  heap = _launch_trace(launch, d_desc, n_jobs, int(nfg.sum()), int(nfg.max()) if n_jobs else 0, paths, out_len, out_np,
                       out_status, out_stats, counter)
  keep = (ws, pdrf, claim, keys, hist, cursor, scratch, d_desc, d_targets, counter, heap)   # alive until the kernel is done
  return dict(desc=desc, paths=paths, out_len=out_len, out_np=out_np, out_status=out_status, out_stats=out_stats,
              d_dbf=d_dbf, n_jobs=n_jobs, timings=timings, tmark=tmark, keep=keep,
              again=(d_cc, d_dbf, shape, anisotropy, jobs, params, n_rows, timings), pool_scale=int(pool_scale))


# the path loop of a label above TRACE_TEAM_MIN voxels runs on a thread-block cluster (at most TRACE_TEAM_MAX of them)
TRACE_TEAM_MIN = int(os.environ.get("B2T_TRACE_TEAM_MIN", 100000))
TRACE_TEAM_MAX = int(os.environ.get("B2T_TRACE_TEAM_MAX", 24))


def _launch_trace(la, d_desc, n_jobs, sum_nfg, max_nfg, paths, out_len, out_np, out_status, out_stats, counter):
  """b2t_trace_batch with the invalidation mode of _lib.invalidation_mode(); returns the buffers the kernel needs alive
  (the strict mode's heap, the teams' state)."""
  L = lib()
  mode, window = _lib.invalidation_mode()
  n_team = min(int((la["nfg_sorted"] >= TRACE_TEAM_MIN).sum()), TRACE_TEAM_MAX, n_jobs)
  team = torch.zeros(max(n_team, 1) * int(L.b2t_trace_team_bytes()), dtype=torch.uint8, device=paths.device)
  heap, heap_words, heap_static = None, 0, 0
  if mode == "strict":
    # static regions (4 entries per voxel) + a spill arena that can take the 27-entries-per-voxel worst case of the
    # largest label and a few dozen average ones besides
    heap_static = int(L.b2t_trace_heap_words(c_u64(sum_nfg), c_u64(n_jobs)))
    spill = 3 * (27 * max_nfg + 2 * max_nfg + 64) + 3 * 27 * min(sum_nfg, 8 * 2 ** 20)
    heap_words = heap_static + spill
    heap = torch.empty(heap_words, dtype=torch.int32, device=paths.device)
  sx, sy, sz = la["shape"]
  an, params = la["anisotropy"], la["params"]
  check(L.b2t_trace_batch(_p(la["d_cc"]), _p(la["d_dbf"]), _p(la["pdrf"]), _p(la["ws"].dist), _p(la["claim"]),
                          _p(la["ws"].stamp), c_i64(sx), c_i64(sy),
                          c_i64(sz), c_f32(an[0]), c_f32(an[1]), c_f32(an[2]), _p(d_desc),
                          c_int(n_jobs), c_f32(params["scale"]), c_f32(params["const"]),
                          c_f32(params["soma_invalidation_scale"]), c_f32(params["soma_invalidation_const"]),
                          c_int(1 if la["fix_branching"] else 0), c_int(NBUCKETS), _p(la["keys"]), _p(la["hist"]),
                          _p(la["cursor"]), _p(la["scratch"]), _p(paths), _p(la["d_targets"]),
                          _p(out_len), _p(out_np), _p(out_status), _p(out_stats), _p(counter),
                          c_int(_lib._MODES[mode]), c_f32(window), _p(heap), c_u64(heap_words), c_u64(heap_static),
                          c_int(n_team), _p(team), stream_ptr()),
        "b2t_trace_batch")
  return heap, team


def trace_arena_finish(st):
  """Wait for the path-loop kernel, compact the path pool, fetch radii (trace.py:186-187)."""
  L = lib()
  desc, n_jobs, timings = st["desc"], st["n_jobs"], st["timings"]
  tmark = st["tmark"]
  dev = st["paths"].device

  def lap(name):
    nonlocal tmark
    if timings is not None:
      torch.cuda.synchronize()
      now = time.perf_counter()
      timings[name] = timings.get(name, 0.0) + (now - tmark)
      tmark = now

  h_len = st["out_len"].cpu().numpy().astype(np.int64)
  h_status = st["out_status"].cpu().numpy()
  lap("paths")
  if (h_status != 0).any():
    bad = int(np.flatnonzero(h_status != 0)[0])
    params = st["again"][5]
    if (h_status[h_status != 0] == -4).all() and not params.get("fix_branching", True) and st["pool_scale"] < 4096:
      # fix_branching=False: root -> target walks outgrew the path pool (no a-priori bound, see trace_arena_start);
      # trace the arena again with a pool 8x the size.  Roots are already in the job table.
      st["keep"] = None
      st["paths"] = None
      return trace_arena_finish(trace_arena_start(*st["again"], pool_scale=8 * st["pool_scale"]))
    what = ("a buffer was too small (B2T_ERR_CAPACITY: path pool, or the strict mode's heap arena)"
            if int(h_status[bad]) == -4 else "an internal error")
    raise B2TError(f"trace kernel reported status {int(h_status[bad])} for cc label {int(desc[bad]['segid'])}: {what}")
  seg_off = np.zeros(n_jobs + 1, dtype=np.int64)
  np.cumsum(h_len, out=seg_off[1:])
  total = int(seg_off[-1])
  d_vox = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
  d_rad = torch.empty(max(total, 1), dtype=torch.float32, device=dev)
  d_srcoff = _dev(desc["path_off"].astype(np.uint32).view(np.int32))
  d_dstoff = _dev(seg_off[:-1].astype(np.uint64).view(np.int64))
  check(L.b2t_gather_paths(_p(st["paths"]), _p(d_srcoff), _p(st["out_len"]), _p(d_dstoff), c_u32(n_jobs),
                           _p(st["d_dbf"]), _p(d_vox), _p(d_rad), stream_ptr()), "b2t_gather_paths")
  lap("gather")
  stats = {"stats": st["out_stats"].cpu().numpy().reshape(-1, 4), "npaths": st["out_np"].cpu().numpy(),
           "segids": desc["segid"].copy()}
  st["keep"] = None
  return d_vox[:total], d_rad[:total], seg_off, desc["segid"].astype(np.int64), stats


def trace_arena(d_cc, d_dbf, shape, anisotropy, jobs, params, n_rows, timings=None):
  """Root, DAF, PDRF, path loop, path buffers for every label of one arena (kimimaro/trace.py:36-194).
  Returns (vox i32 device [N] with -1 path terminators, radii f32 device [N], seg_off int64 [n_jobs+1], seg ids, stats)."""
  return trace_arena_finish(trace_arena_start(d_cc, d_dbf, shape, anisotropy, jobs, params, n_rows, timings))


# ------------------------------------------------------------------------------------------------
# skeleton assembly (trace.py:182-192, intake.py:509-517, 587-593) for all labels at once
# ------------------------------------------------------------------------------------------------
def assemble(d_vox, d_rad, seg_off, seg_ids, shape, anisotropy, offset=(0, 0, 0), group_ids=None, seg_private=None):
  """
  Skeleton.from_path / simple_merge / consolidate (trace.py:182-184) for ALL labels at once on the device:
  unique vertices per label in lexicographic (x,y,z) order, edges remapped / sorted / unique / no self loops, vertices
  without edges dropped, radii of the vertex (first occurrence), then voxel -> physical units in float32 exactly like
  intake.py:509-513.  group_ids (one id per segment, default: the segment's own cc id) lets several connected components
  of one original label be consolidated together, which is what intake.py:587-593 (merge) does afterwards: components
  are disjoint voxel sets, so merging is the same sort over the union.
  One CTA per group (b2t_assemble: a dense u32 volume as the hash set, two shared-memory bitonic sorts); a group above
  b2t_assemble_group_cap() path entries -- a handful of giant labels -- goes through _assemble_general (torch sort /
  unique as plumbing).  seg_private marks segments traced in a private arena: their paths may run through filled voids,
  i.e. voxels of OTHER labels, so a group holding any is assembled in a launch of its own.
  Returns {group id: (vertices f32 [N,3] physical, edges u32 [M,2], radii f32 [N])} (numpy views).
  """
  n = int(d_vox.numel())
  if n == 0:
    return {}
  L = lib()
  if not hasattr(L, "b2t_assemble") or os.environ.get("B2T_ASSEMBLE", "1") == "0":
    return _assemble_general(d_vox, d_rad, seg_off, seg_ids, shape, anisotropy, offset, group_ids)
  sx, sy, sz = shape
  V = sx * sy * sz
  dev = d_vox.device
  seg_off = np.asarray(seg_off, dtype=np.int64)
  n_seg = int(seg_off.size - 1)
  if group_ids is None:
    group_ids = seg_ids
  groups, grank = np.unique(np.asarray(group_ids), return_inverse=True)
  n_grp = int(groups.size)
  order = np.argsort(grank, kind="stable")                   # segments of a group consecutive, path order kept inside
  lens = np.diff(seg_off)
  seg_start = seg_off[:-1][order]
  seg_len = lens[order]
  grp_seg = np.searchsorted(grank[order], np.arange(n_grp + 1))
  ent = np.zeros(n_grp, dtype=np.int64)
  np.add.at(ent, grank, lens)
  grp_out = np.concatenate(([0], np.cumsum(ent)))
  cap = int(L.b2t_assemble_group_cap())
  priv = np.zeros(n_grp, dtype=bool)
  if seg_private is not None and np.any(seg_private):
    priv[np.unique(grank[np.asarray(seg_private, dtype=bool)])] = True
  small = ent <= cap
  launches = [np.flatnonzero(small & ~priv)] + [np.array([g]) for g in np.flatnonzero(small & priv)]
  tab = np.concatenate([seg_start, seg_len, grp_seg, grp_out] + launches).astype(np.uint32)
  d_tab = _dev(tab.view(np.int32))
  o_ss, o_sl, o_gs, o_go, o_l = (int(v) * 4 for v in np.cumsum([0, n_seg, n_seg, n_grp + 1, n_grp + 1]))
  stamp = torch.full((V,), -1, dtype=torch.int32, device=dev)
  out_f = torch.empty(4 * n, dtype=torch.float32, device=dev)          # vertices [3n] | radii [n]
  out_u = torch.zeros(2 * n + 2 * n_grp, dtype=torch.int32, device=dev)  # edges [2n] | counts [2 * n_grp]
  an = [float(a) for a in anisotropy]
  base = int(d_tab.data_ptr())
  lo = 0
  for lst in launches:
    if lst.size:
      check(L.b2t_assemble(_p(d_vox), _p(d_rad), c_vp(base + o_ss), c_vp(base + o_sl), c_vp(base + o_gs), c_vp(base + o_go),
                           c_vp(base + o_l + 4 * lo), c_u32(int(lst.size)), _p(stamp), c_i64(sx), c_i64(sy), c_i64(sz),
                           c_f32(an[0]), c_f32(an[1]), c_f32(an[2]), c_f32(float(offset[0])), c_f32(float(offset[1])),
                           c_f32(float(offset[2])), _p(out_f), c_vp(out_f.data_ptr() + 12 * n), _p(out_u),
                           c_vp(out_u.data_ptr() + 8 * n), stream_ptr()), "b2t_assemble")
    lo += int(lst.size)
  h_f = out_f.cpu().numpy()
  h_u = out_u.cpu().numpy().view(np.uint32)
  h_verts, h_rad = h_f[:3 * n].reshape(-1, 3), h_f[3 * n:]
  h_edges, h_cnt = h_u[:2 * n].reshape(-1, 2), h_u[2 * n:].reshape(-1, 2)
  out = {}
  gl, offs, cnts = groups.tolist(), grp_out.tolist(), h_cnt.tolist()
  for k in np.flatnonzero(small).tolist():
    nv, ne = cnts[k]
    if nv == 0:
      continue
    o = offs[k]
    out[gl[k]] = (h_verts[o:o + nv], h_edges[o:o + ne], h_rad[o:o + nv])
  big = np.flatnonzero(~small)
  if big.size:                                                # the few giant groups: general path on their segments only
    sel = np.flatnonzero(np.isin(grank, big))
    parts_v = [d_vox[int(seg_off[i]):int(seg_off[i + 1])] for i in sel]
    parts_r = [d_rad[int(seg_off[i]):int(seg_off[i + 1])] for i in sel]
    sub_off = np.concatenate(([0], np.cumsum(lens[sel])))
    out.update(_assemble_general(torch.cat(parts_v), torch.cat(parts_r), sub_off, np.asarray(seg_ids)[sel], shape, anisotropy,
                                 offset, np.asarray(group_ids)[sel]))
  return out


def _assemble_general(d_vox, d_rad, seg_off, seg_ids, shape, anisotropy, offset=(0, 0, 0), group_ids=None):
  """
  Skeleton.from_path / simple_merge / consolidate (trace.py:182-184) for ALL labels at once,
  on the device, with torch sort/unique/cumsum as plumbing (the reference's counterpart is np.unique on
  the host): unique vertices per label in lexicographic (x,y,z) order, edges remapped / sorted / unique /
  no self loops, vertices without edges dropped, radii of the vertex, then voxel -> physical units in
  float32 exactly like intake.py:509-513.
  group_ids (one id per segment, default: the segment's own cc id) lets several connected components of
  one original label be consolidated together, which is what intake.py:587-593 (merge) does afterwards:
  components are disjoint voxel sets, so merging is the same sort over the union.
  Returns {group id: (vertices f32 [N,3] physical, edges u32 [M,2], radii f32 [N])} (numpy views).
  """
  sx, sy, sz = shape
  V = sx * sy * sz
  out = {}
  n = int(d_vox.numel())
  if n == 0:
    return out
  dev = d_vox.device
  n_seg = int(seg_ids.size)
  if group_ids is None:
    group_ids = seg_ids
  groups, group_rank = np.unique(np.asarray(group_ids), return_inverse=True)
  n_grp = int(groups.size)
  lens = torch.as_tensor(np.diff(seg_off), device=dev)
  lab_of = torch.repeat_interleave(torch.as_tensor(group_rank.astype(np.int64), device=dev), lens)
  is_vtx = d_vox != -1                                       # 0xffffffff terminators
  v = d_vox.to(torch.int64) & 0xFFFFFFFF
  z = v // (sx * sy)
  r = v - z * (sx * sy)
  y = r // sx
  x = r - y * sx
  gkey = lab_of * V + ((x * sy + y) * sz + z)                # label-major, then lexicographic (x,y,z)
  vi = torch.nonzero(is_vtx).view(-1)
  uniq, inverse = torch.unique(gkey[vi], sorted=True, return_inverse=True)
  nu = int(uniq.numel())
  uid = torch.full((n,), -1, dtype=torch.int64, device=dev)
  uid[vi] = inverse
  a = torch.nonzero(is_vtx[:-1] & is_vtx[1:]).view(-1)       # consecutive entries of one path
  e0, e1 = uid[a], uid[a + 1]
  lo, hi = torch.minimum(e0, e1), torch.maximum(e0, e1)
  keep = lo != hi
  ekey = torch.unique(lo[keep] * nu + hi[keep], sorted=True)
  ea, eb = ekey // nu, ekey % nu
  used = torch.zeros(nu, dtype=torch.bool, device=dev)
  used[ea] = True
  used[eb] = True
  newid = torch.cumsum(used, 0) - 1
  # radius of a vertex = the DBF at its FIRST occurrence (consolidate keeps the first).  Within one arena every
  # occurrence carries the same value; a vertex that a private arena (its own EDT after fill_voids) shares with another
  # component of the same label does not, and a plain scatter would pick one at random.
  first = torch.full((nu,), n, dtype=torch.int64, device=dev)
  first.scatter_reduce_(0, inverse, vi, reduce="amin")
  urad = d_rad[first]
  ulab = uniq // V
  ck = uniq - ulab * V
  ux = ck // (sy * sz)
  uy = (ck // sz) % sy
  uz = ck % sz
  ulab_k = ulab[used]
  verts = torch.stack([ux[used], uy[used], uz[used]], dim=1).to(torch.float32)
  an = torch.as_tensor(np.asarray(anisotropy, dtype=np.float32), device=dev)
  off = torch.as_tensor(np.asarray(offset, dtype=np.float32), device=dev)
  verts = (verts + off) * an                                 # float32, intake.py:509-513
  ea, eb = newid[ea], newid[eb]
  elab = ulab_k[ea]
  bounds = torch.arange(n_grp + 1, device=dev, dtype=torch.int64)
  vstart = torch.searchsorted(ulab_k.contiguous(), bounds)
  estart = torch.searchsorted(elab.contiguous(), bounds)
  edges = torch.stack([ea - vstart[elab], eb - vstart[elab]], dim=1).to(torch.int32)
  h_verts = verts.cpu().numpy()
  h_edges = edges.cpu().numpy().view(np.uint32)
  h_rad = urad[used].cpu().numpy()
  h_vs = vstart.cpu().numpy()
  h_es = estart.cpu().numpy()
  h_vs, h_es, gl = h_vs.tolist(), h_es.tolist(), groups.tolist()
  for k in range(n_grp):
    v0, v1 = h_vs[k], h_vs[k + 1]
    if v1 == v0:
      continue
    out[gl[k]] = (h_verts[v0:v1], h_edges[h_es[k]:h_es[k + 1]], h_rad[v0:v1])
  return out
