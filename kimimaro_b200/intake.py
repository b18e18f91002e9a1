"""
kimimaro_b200.skeletonize -- drop-in for kimimaro.skeletonize (kimimaro/intake.py:58-143) with the
per-label TEASAR hot path on a B200.  Same signature, same teasar_params dictionary, same result
type ({segid: Skeleton}); no CPU fallback.

Deviations from the reference, all explicit:
  * parallel      one process drives one GPU and the device traces all labels concurrently, so the
                  argument is accepted and ignored here; multi-GPU runs shard labels across
                  processes (kimimaro_b200.distributed).
  * voxel_graph, CrackleArray input: not built yet
                   -> NotImplementedError, never a silent CPU path.
                  fill_holes and fix_avocados are built (fill_all_holes, engage_avocado_protection below:
                  b2t_fill_voids per component, b2t_edt_ws after every pass that changed something).
  * tie rules T1-T5 (oracle/oracle.c header) where the reference leaves ties to heap / sort internals.
"""
import ctypes
import functools
import gc
import os
import time
from collections import defaultdict

import numpy as np
import torch

from . import _lib, engine
from ._lib import B2TError, check, lib, stream_ptr  # noqa: F401  (B2TError is part of the module's surface)
from .ops import edt
from .skeleton import Skeleton

c_vp, c_i64, c_u64, c_int = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int
_lib.declare("b2t_fill_voids", [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_u64, c_vp, c_vp])


class DimensionError(Exception):
  pass


DEFAULT_TEASAR_PARAMS = {       # kimimaro/intake.py:47-56
  "scale": 1.5,
  "const": 300,
  "pdrf_scale": 100000,
  "pdrf_exponent": 4,
  "soma_acceptance_threshold": 3500,
  "soma_detection_threshold": 750,
  "soma_invalidation_const": 300,
  "soma_invalidation_scale": 2,
}

_VIEW = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}
_TVIEW = {1: torch.uint8, 2: torch.int16, 4: torch.int32, 8: torch.int64}
_NP_OF_TORCH = {torch.uint8: np.uint8, torch.int8: np.int8, torch.int16: np.int16, torch.int32: np.int32,
                torch.int64: np.int64, torch.uint16: np.uint16, torch.uint32: np.uint32, torch.uint64: np.uint64}


def format_labels(labels, in_place, keep_c_order=False):
  """kimimaro/intake.py:315-342.  keep_c_order: a C-contiguous array is returned as it is (the caller uploads it and
  transposes on the device: a transposing host copy of a 512^3 volume takes several hundred ms)."""
  labels = np.asarray(labels)
  # The reference copies unless in_place (it edits the array: fastremap.renumber, masking).  Here the host array is only
  # ever READ -- it is uploaded, all edits happen on the device copy -- so a Fortran-contiguous array is used as it is
  # whatever in_place says (the copy of a 512^3 uint32 volume costs 100 ms, more than the whole pass).
  if not labels.flags["F_CONTIGUOUS"] and not (keep_c_order and labels.flags["C_CONTIGUOUS"]):
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
  if labels.dtype.kind not in "iu":
    raise TypeError("labels must be of integer or boolean type, got {}".format(labels.dtype))
  return labels


_STAGE = {}          # one pinned staging buffer, kept between calls (pinning memory costs more than a pass)
_STAGE_CHUNK = 32 << 20
_STAGE_MIN = 64 << 20


_CUDART = None


def _host_ptr_is_pinned(ptr):
  """cudaPointerGetAttributes on a host pointer: page-locked (cudaMemoryTypeHost) or not.  (Tensor.is_pinned() of a
  tensor made by torch.from_numpy over a view of pinned memory answered False on the GPU box -- call 59 -- and sent
  bench.py's pinned buffer through the staging path: 14 instead of 9.7 ms.)"""
  global _CUDART
  if _CUDART is None:
    _CUDART = False
    for cand in ("libcudart.so.12", "libcudart.so", "/usr/local/cuda/lib64/libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
      try:
        _CUDART = ctypes.CDLL(cand)
        break
      except OSError:
        pass
  if not _CUDART:
    return None

  class _Attr(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int), ("device", ctypes.c_int), ("devicePointer", ctypes.c_void_p),
                ("hostPointer", ctypes.c_void_p), ("pad", ctypes.c_byte * 64)]
  a = _Attr()
  if _CUDART.cudaPointerGetAttributes(ctypes.byref(a), ctypes.c_void_p(ptr)) != 0:
    _CUDART.cudaGetLastError()
    return None
  return a.type == 1                                           # cudaMemoryTypeHost


def _upload(flat):
  """Host -> device copy of the label volume.  A pinned array goes down in one asynchronous copy (9.5 ms per 512 MiB).
  An ordinary (pageable) numpy array would crawl through the driver's small bounce buffer (150 ms per 512 MiB measured):
  it is copied chunk by chunk into a pinned staging buffer by a few threads (numpy releases the GIL for the copy) and
  every chunk is sent on as soon as it has landed, so the host copy and the upload overlap."""
  t = torch.from_numpy(flat)
  nbytes = flat.nbytes
  if nbytes < _STAGE_MIN or not torch.cuda.is_available() or os.environ.get("B2T_STAGED_UPLOAD", "1") == "0":
    return t.cuda(non_blocking=True)
  pinned = _host_ptr_is_pinned(flat.ctypes.data)
  if pinned is None:
    pinned = t.is_pinned()
  if pinned:
    return t.cuda(non_blocking=True)
  from concurrent.futures import ThreadPoolExecutor
  buf = _STAGE.get("buf")
  if buf is None or buf.numel() < nbytes:
    _STAGE.clear()
    buf = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    _STAGE["buf"] = buf
    _STAGE["pool"] = ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1))
  pool = _STAGE["pool"]
  src = flat.view(np.uint8)
  stage = buf.numpy()
  d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
  bounds = list(range(0, nbytes, _STAGE_CHUNK)) + [nbytes]
  futs = [pool.submit(np.copyto, stage[a:b], src[a:b]) for a, b in zip(bounds[:-1], bounds[1:])]
  for f, a, b in zip(futs, bounds[:-1], bounds[1:]):
    f.result()
    d[a:b].copy_(buf[a:b], non_blocking=True)
  torch.cuda.current_stream().synchronize()                   # the staging buffer is reused by the next call
  return d.view(t.dtype)


def _merge_params(teasar_params):
  params = dict(engine.TRACE_DEFAULTS)
  for k, v in teasar_params.items():
    if k not in params:
      # trace(**teasar_params) raises TypeError on unknown keywords (intake.py:503)
      raise TypeError("trace() got an unexpected keyword argument '{}'".format(k))
    params[k] = v
  return params


def _find_soma_root(d_dbf, shape, dbf_max):
  """trace.py:269-289 on the device-resident DBF of a private arena."""
  sx, sy, sz = shape
  idx = torch.nonzero(d_dbf == float(dbf_max)).view(-1).cpu().numpy().astype(np.int64)
  z = idx // (sx * sy)
  r = idx - z * (sx * sy)
  y = r // sx
  x = r - y * sx
  coords = np.stack([x, y, z], axis=1)
  # np.where order is C order over (x,y,z): sort accordingly so that argmin's first-minimum matches
  order = np.lexsort((coords[:, 2], coords[:, 1], coords[:, 0]))
  coords = coords[order]
  com = np.asarray(coords.astype(np.float64).sum(axis=0) / float(coords.shape[0]), dtype=np.float32)
  root = np.argmin(np.sum((coords - com) ** 2, axis=1))
  return tuple(coords[root].astype(np.uint32))


def synapses_to_targets(labels, synapses, progress=False):
  """synapses_to_targets (kimimaro/intake.py:706-745): for every label and every swc_label of its synapses, the
  label's voxel nearest to each centroid (voxel units, same origin as labels) becomes a skeletonization target.
  labels: 3-D array; synapses: { label: [ (centroid, swc_label), ... ] }.  Returns { (x,y,z): swc_label }.
  Host-side preparation of extra_targets_*: numpy only, like the reference (no kernel involved)."""
  labels = np.asarray(labels)
  while labels.ndim > 3:
    labels = labels[..., 0]
  targets = {}
  for label, pairs in synapses.items():
    point_cloud = np.vstack((labels == label).nonzero()).T     # [ [x,y,z], ... ] in C order, like the reference
    if len(point_cloud) == 0:
      continue
    swc_labels = defaultdict(list)
    for centroid, swc_label in pairs:
      swc_labels[swc_label].append(centroid)
    for swc_label, centroids in swc_labels.items():
      cents = np.asarray(centroids, dtype=np.float64).reshape(len(centroids), -1)
      pc = point_cloud.astype(np.float64)
      # scipy.spatial.distance.cdist(point_cloud, centroids): Euclidean, float64; argmin takes the first minimum
      distances = np.sqrt(((pc[:, None, :] - cents[None, :, :]) ** 2).sum(axis=2))
      minima = np.unique(np.argmin(distances, axis=0))
      targets.update({tuple(int(v) for v in point_cloud[idx]): swc_label for idx in minima})
  return targets


def _fill_voids(mask, shape):
  """fill_voids.fill(mask, in_place=True, return_fill_count=True) (trace.py:109, intake.py:779) on a flat uint8
  device mask in Fortran order; returns the number of voxels filled."""
  V = shape[0] * shape[1] * shape[2]
  dev = mask.device
  reach = torch.empty(V, dtype=torch.int32, device=dev)
  queue = torch.empty(2 * V, dtype=torch.int32, device=dev)
  ctrl = torch.zeros(16, dtype=torch.int32, device=dev)
  check(lib().b2t_fill_voids(c_vp(mask.data_ptr()), c_i64(shape[0]), c_i64(shape[1]), c_i64(shape[2]),
                             c_vp(reach.data_ptr()), c_vp(queue.data_ptr()), c_u64(V), c_vp(ctrl.data_ptr()),
                             stream_ptr()), "b2t_fill_voids")
  return int(ctrl[5].item())


def fill_all_holes(d_cc, shape, n_cc, h_count, h_bbox, fill_fn=None, return_fill_count=False):
  """fill_all_holes (kimimaro/intake.py:747-794) on the device-resident cc volume, in place: every connected
  component, in label order, gets its voids filled (b2t_fill_voids on its bounding-box crop); components that end up
  inside an earlier one disappear and are not visited.  h_count / h_bbox: per-label voxel counts and inclusive
  bounding boxes (x0, y0, z0, x1, y1, z1) of the unfilled volume."""
  fill_fn = fill_fn or _fill_voids
  sx, sy, sz = shape
  d_cc3 = d_cc.view(sz, sy, sx)
  alive = np.ones(n_cc + 1, dtype=bool)
  pixels_filled = 0
  for label in range(1, n_cc + 1):
    if not alive[label] or h_count[label] == 0:
      continue
    x0, y0, z0, x1, y1, z1 = (int(v) for v in h_bbox[label])
    cshape = (x1 - x0 + 1, y1 - y0 + 1, z1 - z0 + 1)
    if cshape[0] * cshape[1] * cshape[2] == int(h_count[label]):
      continue                                                # the component is its whole box: nothing to fill
    crop = d_cc3[z0:z1 + 1, y0:y1 + 1, x0:x1 + 1]
    mask = (crop == label).to(torch.uint8).contiguous().view(-1)
    n = fill_fn(mask, cshape)
    pixels_filled += n
    if n == 0:
      continue
    m3 = mask.view(cshape[2], cshape[1], cshape[0]) != 0
    for sub in torch.unique(crop[m3]).cpu().tolist():         # the components that were swallowed
      if sub != label and sub != 0:
        alive[sub] = False
    crop[m3] = label                                          # writes through to d_cc
  if return_fill_count:
    return d_cc, pixels_filled
  return d_cc


def _fill_voids_2d(plane, fill_fn):
  """fill_voids.fill on a 2-D image (paint_walls, intake.py:666-677) with the 3-D kernel: the image is the middle
  slice of a three-slice volume whose outer slices are solid, so its background connects only within the slice and
  only the image's own border pixels lie on a face of the volume."""
  b, a = plane.shape
  if a == 0 or b == 0:
    return plane
  vol = torch.ones((3, b, a), dtype=torch.uint8, device=plane.device)
  vol[1] = plane.to(torch.uint8)
  flat = vol.view(-1)
  fill_fn(flat, (a, b, 3))
  return flat.view(3, b, a)[1] != 0


def find_avocado_fruit(xline, yline, zline, cx, cy, cz, background=0):
  """kimimaro.skeletontricks.find_avocado_fruit (pyx:905-992) on the three axis lines of the label volume through
  (cx, cy, cz) (host arrays): the first foreign label on each of the six rays votes; a ray that reaches background
  first, or the array end, does not.  The rays towards smaller coordinates stop before index 0, like the reference.
  Returns (pit, fruit)."""
  label = int(xline[cx])
  changes = []
  for line, c in ((xline, cx), (yline, cy), (zline, cz)):
    for rng in (range(c, len(line)), range(c, 0, -1)):
      for i in rng:
        v = int(line[i])
        if v == background:
          break
        if v != label:
          changes.append(v)
          break
  if len(changes) < 3:
    return (label, label)
  allowed_differences = 1 if len(changes) > 3 else 0
  uniq, cts = np.unique(changes, return_counts=True)
  k = int(np.argmax(cts))
  if len(changes) - int(cts[k]) > allowed_differences:
    return (label, label)
  return (label, int(uniq[k]))


def _default_stats(d_cc, d_dbf, shape, n_cc):
  count, bbox, _, _ = engine.label_stats(d_cc, d_dbf, shape, n_cc)
  return count.cpu().numpy(), bbox.cpu().numpy().reshape(-1, 6)


# The six faces of a [z, y, x] crop in the order the reference paints them (paint_walls, intake.py:666-677):
# z faces, then y faces, then x faces.
_WALLS = tuple((axis, end) for axis in (0, 1, 2) for end in (0, -1))


def _wall(img, axis, end):
  sl = [slice(None)] * 3
  sl[axis] = end
  return tuple(sl)


def _sealed_pit_image(crop, label, fill_fn):
  """Mask of `label` in its bounding-box crop with every face closed by a 2-D fill, so that a pit which touches the
  box's faces still encloses its interior for the 3-D fill that follows."""
  img = (crop == label)
  for axis, end in _WALLS:
    sl = _wall(img, axis, end)
    img[sl] = _fill_voids_2d(img[sl], fill_fn)
  return img


def _deepest_voxel(img, dbf_crop, origin):
  """Voxel of the mask with the largest DBF, first one in Fortran raster order (intake.py:595-598); volume coordinates."""
  ez, ey, ex = img.shape
  k = int(torch.argmax((img.to(torch.float32) * dbf_crop).reshape(-1)).item())
  kz, rem = divmod(k, ex * ey)
  ky, kx = divmod(rem, ex)
  return kx + origin[0], ky + origin[1], kz + origin[2]


class _PitLedger:
  """Which labels a pass left alone and which it merged (the two sets of intake.py:646-704).  A label that has taken
  part in a merge can no longer count as left alone, whichever side of the merge it was on."""
  def __init__(self):
    self.settled, self.merged = set(), set()

  def record(self, pit, fruit):
    """True when `pit` is absorbed by `fruit`."""
    if pit == fruit and pit not in self.merged:
      self.settled.add(pit)
      return False
    self.settled -= {pit, fruit}
    self.merged |= {pit, fruit}
    return True


def engage_avocado_protection_single_pass(d_cc, d_dbf, shape, n_cc, candidates, stats_fn, fill_fn):
  """One pass of the avocado protection (intake.py:646-704) on the device-resident volumes; `candidates` come in the
  reference's iteration order.  Returns (labels left alone, labels that took part in a merge)."""
  book = _PitLedger()
  todo = [c for c in candidates if c != 0]
  if not todo:
    return book.settled, book.merged
  sx, sy, sz = shape
  d_cc3 = d_cc.view(sz, sy, sx)
  d_dbf3 = d_dbf.view(sz, sy, sx)
  _, h_bbox = stats_fn(d_cc, d_dbf, shape, n_cc)              # boxes as they are when the pass starts (intake.py:679)
  for pit_label in todo:
    x0, y0, z0, x1, y1, z1 = (int(v) for v in h_bbox[pit_label])
    box = (slice(z0, z1 + 1), slice(y0, y1 + 1), slice(x0, x1 + 1))
    crop = d_cc3[box]
    img = _sealed_pit_image(crop, pit_label, fill_fn)
    cx, cy, cz = _deepest_voxel(img, d_dbf3[box], (x0, y0, z0))
    pit, fruit = find_avocado_fruit(d_cc3[cz, cy, :].cpu().numpy(), d_cc3[cz, :, cx].cpu().numpy(),
                                    d_cc3[:, cy, cx].cpu().numpy(), cx, cy, cz)
    if book.record(pit, fruit):
      img |= (crop == fruit)
    flat = img.to(torch.uint8).contiguous().view(-1)
    fill_fn(flat, (x1 - x0 + 1, y1 - y0 + 1, z1 - z0 + 1))
    crop[flat.view(img.shape) != 0] = fruit                    # writes through to d_cc
  return book.settled, book.merged


def _hot_labels(d_cc, hot):
  """set(fastremap.unique(cc_labels * (all_dbf > t))) (intake.py:617) without forming the product.  The reference then
  ITERATES this Python set, and a set's order depends on how it was filled (table growth, tombstones), so the set is
  built by the same sequence of insertions: the sorted unique values of the product, 0 included when the product has a
  zero anywhere.  The caller removes entries in place, like the reference's `-=` and discard."""
  seen = torch.unique(d_cc[hot]).cpu().tolist()
  zero_somewhere = bool(((d_cc == 0) | ~hot).any().item())
  out = set()
  for v in ([0] if zero_somewhere else []) + [int(v) for v in seen if v != 0]:
    out.add(v)
  return out


def _last_run_owner(d_cc, d_before, n_cc):
  """get_mapping(orig_cc_labels, cc_labels) (pyx:490-525) for the surviving labels: the START of the last run of a label
  in raster order names its pre-protection component; -1 for labels that are gone."""
  starts = torch.nonzero(d_cc[1:] != d_cc[:-1]).view(-1) + 1
  starts = torch.cat([torch.zeros(1, dtype=starts.dtype, device=starts.device), starts])
  last = torch.full((n_cc + 1,), -1, dtype=torch.int64, device=d_cc.device)
  last.scatter_reduce_(0, d_cc[starts].to(torch.int64), starts, reduce="amax")
  h_last = last.cpu().numpy()
  h_src = np.full(n_cc + 1, -1, dtype=np.int64)
  ok = h_last >= 0
  h_src[ok] = d_before[torch.as_tensor(h_last[ok], device=d_cc.device)].cpu().numpy().astype(np.int64)
  return h_src


def engage_avocado_protection(d_cc, d_dbf, shape, n_cc, soma_detection_threshold, edtfn, stats_fn=None, fill_fn=None,
                              max_passes=20):
  """engage_avocado_protection (kimimaro/intake.py:600-644): a nucleus segmented apart from its cell -- the pit of an
  avocado -- takes the label of the fruit around it; up to 20 passes for nested cases, the EDT redone after every pass
  that changed something.  d_cc is edited in place.  Instead of renumbering (intake.py:636, not visible in the result)
  the surviving labels keep their numbers; returns (d_dbf, last) where last[L] is the pre-protection component that
  get_mapping would name for the surviving component L, -1 for labels that are gone."""
  stats_fn = stats_fn or _default_stats
  fill_fn = fill_fn or _fill_voids
  d_before = d_cc.clone()
  left_alone = set()
  cutoff = soma_detection_threshold / 2.5
  for _ in range(max_passes):
    # labels a previous pass left alone are not looked at again
    todo = _hot_labels(d_cc, d_dbf > cutoff)
    todo -= left_alone
    todo.discard(0)
    settled, merged = engage_avocado_protection_single_pass(d_cc, d_dbf, shape, n_cc, todo, stats_fn, fill_fn)
    left_alone |= settled
    if not merged:
      break
    d_dbf = edtfn(d_cc)
  return d_dbf, _last_run_owner(d_cc, d_before, n_cc)


def _private_arena(d_cc3, d_dbf3, segid, bbox, anisotropy, params, root, targets_before, targets_after,
                   timings):
  """A label whose DBF exceeds soma_detection_threshold (trace.py:108-127): crop, fill its voids,
  redo the EDT on the crop if anything was filled, then trace it in its own arena."""
  x0, y0, z0, x1, y1, z1 = bbox
  shape = (x1 - x0 + 1, y1 - y0 + 1, z1 - z0 + 1)
  V = shape[0] * shape[1] * shape[2]
  dev = d_cc3.device
  crop = d_cc3[z0:z1 + 1, y0:y1 + 1, x0:x1 + 1]
  mask = (crop == segid).to(torch.uint8).contiguous().view(-1)
  filled = _fill_voids(mask, shape)
  if filled > 0:
    dbf = edt(mask, shape, anisotropy, black_border=bool(mask.all().item()))
  else:
    dbf = torch.where(mask.view(shape[2], shape[1], shape[0]) != 0,
                      d_dbf3[z0:z1 + 1, y0:y1 + 1, x0:x1 + 1], torch.zeros((), device=dev)).contiguous().view(-1)
  dbf_max = np.float32(dbf.max().item())
  soma_mode = bool(dbf_max > params["soma_acceptance_threshold"])
  off = np.array([x0, y0, z0], dtype=np.int64)

  def local(pt):
    p = np.asarray(pt, dtype=np.int64) - off
    return int(p[0] + shape[0] * (p[1] + shape[1] * p[2]))

  tb = [local(t) for t in targets_before]
  ta = [local(t) for t in targets_after]
  n_fg = int(mask.sum().item())
  rlin, first, soma_radius, free_space = -1, 0, 0.0, 0.0
  if soma_mode:
    if root is not None:
      tb.insert(0, local(root))                               # trace.py:124-125
    r = _find_soma_root(dbf, shape, dbf_max)
    rlin = int(r[0]) + shape[0] * (int(r[1]) + shape[1] * int(r[2]))
    # soma_radius = dbf_max * soma_invalidation_scale + soma_invalidation_const  (trace.py:127)
    soma_radius = np.float32(dbf_max * params["soma_invalidation_scale"] + params["soma_invalidation_const"])
    free_space = float(dbf[rlin].item())                      # trace.py:134
  elif root is not None:
    rlin = local(root)
  else:
    first = int(torch.nonzero(mask)[0].item())                # first_label (pyx:307-326)
  jobs = engine.Jobs([1], [n_fg], [first], [rlin], [dbf_max], tb={0: tb} if tb else {}, ta={0: ta} if ta else {},
                     soma_mode=[soma_mode], soma_radius=[soma_radius], free_space=[free_space])
  cc1 = mask.to(torch.int32)
  vox, rad, seg_off, seg_ids, stats = engine.trace_arena(cc1, dbf, shape, anisotropy, jobs, params, 1, timings)
  # path voxels as linear indices of the whole volume (terminators stay -1), so that they join the other labels' paths
  full = d_cc3.shape[2], d_cc3.shape[1], d_cc3.shape[0]
  v = vox.to(torch.int64) & 0xFFFFFFFF
  lz = v // (shape[0] * shape[1])
  r = v - lz * (shape[0] * shape[1])
  ly = r // shape[0]
  lx = r - ly * shape[0]
  g = (lx + x0) + full[0] * ((ly + y0) + full[1] * (lz + z0))
  vox_g = torch.where(vox == -1, vox, g.to(torch.int32))
  return (vox_g, rad, np.diff(seg_off)), stats


def _skeletonize(
  all_labels, teasar_params=DEFAULT_TEASAR_PARAMS, anisotropy=(1, 1, 1),
  object_ids=None, dust_threshold=1000,
  progress=True, fix_branching=True, in_place=False,
  fix_borders=True, parallel=1, parallel_chunk_size=100,
  extra_targets_before=[], extra_targets_after=[],
  fill_holes=False, fix_avocados=False,
  voxel_graph=None, timings=None, label_subset=None, device_labels=None, edt_events=None, label_dtype=None, raw_paths=False, border_shard=None,
):
  """
  Skeletonize all non-zero labels in a 2D or 3D image (kimimaro/intake.py:58-143).
  Returns { segid: Skeleton } with vertices in physical units (space='physical').

  Extra keyword arguments (not in the reference): timings (dict filled with per-phase seconds),
  label_subset (callable(list of cc ids) -> list: used by the multi-GPU launcher to shard labels),
  device_labels (a flat Fortran-ordered CUDA tensor already holding the volume: skips the H2D copy;
  all_labels then carries the shape), edt_events (list receiving (start, end) CUDA events around K1).
  """
  if voxel_graph is not None:
    raise NotImplementedError(
      "voxel_graph is not built yet in kimimaro_b200 (SURVEY.md 8f row N4); there is no CPU fallback")
  _lib.require_device()
  params = _merge_params(teasar_params)
  params["fix_branching"] = bool(fix_branching)
  anisotropy = np.array(anisotropy, dtype=np.float32)
  an = tuple(float(a) for a in anisotropy)
  t_all = time.perf_counter()
  tm = timings if timings is not None else None

  def lap(name, t0):
    if tm is not None:
      torch.cuda.synchronize()
      tm[name] = tm.get(name, 0.0) + time.perf_counter() - t0
    return time.perf_counter()

  t0 = time.perf_counter()
  if device_labels is not None:
    shape = tuple(int(s) for s in all_labels)                 # all_labels carries the shape in this mode
    while len(shape) < 3:
      shape = shape + (1,)
    d_labels = device_labels
    size = int(np.prod(shape))
    # ids come back in the dtype the caller's tensor has, or in label_dtype (the host array's dtype when the multi-GPU
    # launcher uploaded it: torch has no kernels for the unsigned types, so the device tensor carries signed bit patterns)
    key_dtype = np.dtype(label_dtype) if label_dtype is not None else _NP_OF_TORCH.get(device_labels.dtype)
  else:
    all_labels = format_labels(all_labels, in_place=in_place, keep_c_order=True)
    shape = all_labels.shape
    size = all_labels.size
    key_dtype = all_labels.dtype
    if size <= dust_threshold:
      return {}
    if all_labels.flags["F_CONTIGUOUS"]:
      flat = all_labels.reshape(-1, order="F")
      d_labels = _upload(flat.view(_VIEW[flat.dtype.itemsize]))
    else:                                                      # C order: upload the bytes as they lie, transpose on the device
      flat = all_labels.reshape(-1)
      d_c = _upload(flat.view(_VIEW[flat.dtype.itemsize]))
      if d_c.dtype in (torch.uint16, torch.uint32, torch.uint64):
        d_c = d_c.view(_TVIEW[d_c.element_size()])
      d_labels = d_c.view(shape[0], shape[1], shape[2]).permute(2, 1, 0).contiguous().view(-1)
      del d_c
  if size <= dust_threshold:
    return {}
  if d_labels.dtype in (torch.uint16, torch.uint32, torch.uint64):
    d_labels = d_labels.view(_TVIEW[d_labels.element_size()])
  sx, sy, sz = shape
  V = size
  t0 = lap("h2d", t0)

  if object_ids is not None:                                   # apply_object_mask (intake.py:519-535)
    ids = torch.as_tensor(np.asarray(object_ids).astype(np.int64), device=d_labels.device).to(d_labels.dtype)
    d_labels = torch.where(torch.isin(d_labels, ids), d_labels, torch.zeros((), dtype=d_labels.dtype,
                                                                           device=d_labels.device))
  nz = d_labels != 0
  if not bool(nz.any().item()):
    return {}
  black_border = bool(nz.all().item()) and bool((d_labels == d_labels[0]).all().item())   # minlabel == maxlabel
  del nz

  # ---- preamble: connected components, EDT, per-label statistics ----
  d_cc, n_cc = engine.connected_components(d_labels, shape)
  t0 = lap("ccl", t0)
  if fill_holes:                                               # intake.py:166-167
    # boxes and counts of the unfilled components; the statistics kernel wants a DBF: any float volume will do
    c0, b0, _, _ = engine.label_stats(d_cc, torch.zeros(V, dtype=torch.float32, device=d_cc.device), shape, n_cc)
    fill_all_holes(d_cc, shape, n_cc, c0.cpu().numpy(), b0.cpu().numpy().reshape(-1, 6))
    t0 = lap("fill_holes", t0)
  if edt_events is not None:
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
  d_dbf = edt(d_cc, shape, an, black_border)
  if edt_events is not None:
    ev1.record()
    edt_events.append((ev0, ev1))
  t0 = lap("edt", t0)
  avocado_src = None
  if fix_avocados:                                             # intake.py:187-193
    _, _, _, first0 = engine.label_stats(d_cc, d_dbf, shape, n_cc)
    h_first0 = first0.cpu().numpy().view(np.uint32).astype(np.int64)
    h_orig0 = d_labels[torch.as_tensor(np.clip(h_first0, 0, V - 1), device=d_labels.device)].cpu().numpy()
    d_dbf, avocado_src = engage_avocado_protection(
      d_cc, d_dbf, shape, n_cc, teasar_params.get("soma_detection_threshold", 0),
      lambda cc: edt(cc, shape, an, black_border))
    t0 = lap("fix_avocados", t0)
  count, bbox, dbfmax, first = engine.label_stats(d_cc, d_dbf, shape, n_cc)
  h_count = count.cpu().numpy()
  h_bbox = bbox.cpu().numpy().reshape(-1, 6)
  h_dbfmax = dbfmax.cpu().numpy()
  h_first = first.cpu().numpy().view(np.uint32).astype(np.int64)
  # remapping: original label at the first voxel of every component (get_mapping, pyx:490-525)
  gather_idx = torch.as_tensor(np.clip(h_first, 0, V - 1), device=d_labels.device)
  h_orig = d_labels[gather_idx].cpu().numpy()
  if avocado_src is not None:                                  # adjusted_remapping (intake.py:637-642)
    ok = avocado_src >= 0
    h_orig = h_orig.copy()
    h_orig[ok] = h_orig0[avocado_src[ok]]
  if key_dtype is not None and h_orig.dtype != key_dtype and h_orig.dtype.itemsize == np.dtype(key_dtype).itemsize:
    h_orig = h_orig.view(key_dtype)                            # signed <-> unsigned bit casts of the upload undone
  t0 = lap("stats", t0)

  cc_segids = [int(s) for s in np.flatnonzero(h_count > dust_threshold) if s != 0]
  if label_subset is not None:
    # cost proxy for the multi-GPU split: voxels, a label that will take the soma branch at half weight (its private
    # arena is throughput kernels, 0.8 ns per voxel on synthetic-512, against 1.6 ns per voxel of ordinary labels)
    cost = h_count.astype(np.float64)
    cost[h_dbfmax > params["soma_detection_threshold"]] *= 0.5
    cc_segids = list(label_subset(cc_segids, cost))

  def points_to_labels(pts):
    mapping = defaultdict(list)
    if len(pts) == 0:
      return mapping
    pa = np.asarray(pts, dtype=np.int64).reshape(-1, 3)
    lin = pa[:, 0] + sx * (pa[:, 1] + sy * pa[:, 2])
    labs = d_cc[torch.as_tensor(lin, device=d_cc.device)].cpu().numpy()
    for pt, l in zip(pa.tolist(), labs.tolist()):
      mapping[int(l)].append(tuple(pt))
    return mapping
  extra_before = points_to_labels(extra_targets_before)
  extra_after = points_to_labels(extra_targets_after)

  # ---- per label arguments of trace() (intake.py:445-504), as one table ----
  segs = np.asarray(cc_segids, dtype=np.int64)
  if segs.size:
    ext = (h_bbox[segs, 3:6].astype(np.int64) - h_bbox[segs, 0:3].astype(np.int64) + 1)
    segs = segs[np.prod(ext, axis=1) > 1]                       # roi.volume() <= 1 is skipped (intake.py:455-456)
  is_private = h_dbfmax[segs] > params["soma_detection_threshold"]   # trace.py:108: takes the fill / re-EDT branch

  border_targets = {}
  early_roots = None
  if fix_borders:
    m0 = segs[~is_private]
    if m0.size and os.environ.get("B2T_EARLY_ROOTS", "0") == "1":
      # find_root for the main arena on a second stream while the border targets are computed (engine.RootSweep).  Off:
      # measured on synthetic-512 (call 31) the sweep's ~2000 CTAs keep every SM slot taken, the border targets' small
      # launches queue behind them and the pass gets slower and erratic (82 -> 92..150 ms)
      early_roots = engine.RootSweep(d_cc, shape, an, h_first[m0], m0, h_count[m0], n_cc)
    # (computing them on a second stream under the volume's EDT and the label statistics was measured: 81.8 vs 81.3 ms
    # per pass on synthetic-512, no gain -- the EDT is over before the first face is queued)
    border_targets = engine.compute_border_targets(d_cc, shape, anisotropy, shard=border_shard)
  t0 = lap("border_targets", t0)
  lin = lambda p: int(p[0]) + sx * (int(p[1]) + sy * int(p[2]))

  def manual_targets(segid):
    tb, ta, root = [], [], None
    bt = border_targets.get(segid)
    if bt is not None and len(bt) > 0:
      tb = [tuple(int(v) for v in p) for p in bt.tolist()]
      root = tb.pop()                                           # intake.py:484-486
    if segid in extra_before:
      tb.extend(extra_before[segid])
    if segid in extra_after:
      ta.extend(extra_after[segid])
    return tb, ta, root

  private = []
  for segid in segs[is_private].tolist():
    tb, ta, root = manual_targets(segid)
    private.append((segid, tuple(int(v) for v in h_bbox[segid]), root, tb, ta))
  main = segs[~is_private]
  roots = np.full(main.size, -1, dtype=np.int64)
  tb_map, ta_map = {}, {}
  has_targets = set(border_targets.keys()) | set(extra_before.keys()) | set(extra_after.keys())
  if has_targets:
    # linear indices of all border targets in one go (the common case: a label with border targets and no extra ones)
    bt_lin = {}
    if border_targets:
      keys = list(border_targets.keys())
      arrs = [border_targets[k] for k in keys]
      allp = np.concatenate(arrs).astype(np.int64).reshape(-1, 3)
      alll = (allp[:, 0] + sx * (allp[:, 1] + sy * allp[:, 2])).tolist()
      o = 0
      for k, a in zip(keys, arrs):
        bt_lin[k] = alll[o:o + len(a)]
        o += len(a)
    for i, segid in enumerate(main.tolist()):
      if segid not in has_targets:
        continue
      if segid in bt_lin and segid not in extra_before and segid not in extra_after:
        l = bt_lin[segid]
        if l:
          roots[i] = l[-1]                                       # the last border target is the root (intake.py:484-486)
          if len(l) > 1:
            tb_map[i] = l[:-1]
        continue
      tb, ta, root = manual_targets(segid)
      if root is not None:
        roots[i] = lin(root)
      if tb:
        tb_map[i] = [lin(p) for p in tb]
      if ta:
        ta_map[i] = [lin(p) for p in ta]
  ws = None
  if early_roots is not None:
    swept = early_roots.roots()
    roots = np.where(roots < 0, swept, roots)
    ws = early_roots.ws
    t0 = lap("find_root", t0)
  jobs = engine.Jobs(main, h_count[main], h_first[main], roots, h_dbfmax[main], tb=tb_map, ta=ta_map,
                     bbox_x=h_bbox[main][:, [0, 3]])

  raw = []                # (path voxels i32 device, radii f32 device, segment lengths, original label per segment)
  stats_all = []
  # The path loop of the main arena is latency-bound (one CTA per label), the private arenas (soma branch)
  # are throughput kernels: run them side by side -- the main path kernel is launched asynchronously with
  # a reduced footprint, the private arenas go to a second stream, then the main arena is collected.
  # Measured on synthetic-512: with the cooperative sweeps capped at 2 blocks/SM the private arena slows down
  # by more than the overlap saves (211 vs 163 ms per pass), so this stays off unless B2T_OVERLAP=1.
  overlap = (os.environ.get("B2T_OVERLAP", "0") == "1") and len(jobs) > 0 and len(private) > 0 and tm is None
  handle = None
  if len(jobs):
    if overlap:
      lib().b2t_set_launch_limits(0, 2)
    handle = engine.trace_arena_start(d_cc, d_dbf, shape, an, jobs, params, n_cc, tm, ws=ws)
    if not overlap:
      vox, rad, seg_off, seg_ids, stats = engine.trace_arena_finish(handle)
      handle = None
      raw.append((vox, rad, np.diff(seg_off), h_orig[seg_ids], np.zeros(seg_ids.size, dtype=bool)))
      stats_all.append(stats)
  if private:
    d_cc3 = d_cc.view(sz, sy, sx)
    d_dbf3 = d_dbf.view(sz, sy, sx)
    main_stream = torch.cuda.current_stream()
    side = torch.cuda.Stream() if overlap else main_stream
    if overlap:
      side.wait_stream(main_stream)
      lib().b2t_set_launch_limits(2, 0)
    try:
      with torch.cuda.stream(side):
        for segid, bb, root, tb, ta in private:
          t0 = time.perf_counter()
          ptm = {} if tm is not None else None
          (pvox, prad, plens), stats = _private_arena(d_cc3, d_dbf3, segid, bb, an, params, root, tb, ta, ptm)
          raw.append((pvox, prad, plens, np.full(plens.size, h_orig[segid], dtype=h_orig.dtype),
                      np.ones(plens.size, dtype=bool)))
          stats_all.append(stats)
          if tm is not None:
            tm["soma"] = tm.get("soma", 0.0) + time.perf_counter() - t0
            for k, v in ptm.items():
              tm["soma_" + k] = tm.get("soma_" + k, 0.0) + v
    finally:
      lib().b2t_set_launch_limits(0, 0)
    if overlap:
      main_stream.wait_stream(side)
  if handle is not None:
    vox, rad, seg_off, seg_ids, stats = engine.trace_arena_finish(handle)
    raw.insert(0, (vox, rad, np.diff(seg_off), h_orig[seg_ids], np.zeros(seg_ids.size, dtype=bool)))
    stats_all.insert(0, stats)
  if tm is not None:
    tm["n_cc"] = n_cc
    tm["n_traced"] = len(jobs) + len(private)
    tm["kernel_stats"] = stats_all
  bundle = join_raw(raw, d_labels.device, h_orig.dtype)
  if raw_paths:
    return bundle
  out = skeletons_from_raw(bundle, shape, anisotropy, tm)
  if tm is not None:
    tm["total"] = time.perf_counter() - t_all
  return out


def join_raw(raw, device, id_dtype):
  """Concatenate per-arena path buffers: (voxels i32 [N] device with -1 terminators, radii f32 [N] device, segment
  lengths int64 [S], original label of every segment [S], traced-in-a-private-arena flag of every segment [S])."""
  if not raw:
    return (torch.empty(0, dtype=torch.int32, device=device), torch.empty(0, dtype=torch.float32, device=device),
            np.zeros(0, dtype=np.int64), np.zeros(0, dtype=id_dtype), np.zeros(0, dtype=bool))
  if len(raw) == 1:
    return raw[0][0], raw[0][1], np.asarray(raw[0][2], dtype=np.int64), np.asarray(raw[0][3]), np.asarray(raw[0][4], dtype=bool)
  return (torch.cat([r[0] for r in raw]), torch.cat([r[1] for r in raw]),
          np.concatenate([np.asarray(r[2], dtype=np.int64) for r in raw]), np.concatenate([np.asarray(r[3]) for r in raw]),
          np.concatenate([np.asarray(r[4], dtype=bool) for r in raw]))


def skeletons_from_raw(bundle, shape, anisotropy, tm=None):
  """Path buffers -> {original label: Skeleton}: one device-side assembly for every label of every arena (the connected
  components of one original label are consolidated together, which is what the reference's merge per id does,
  intake.py:509-517, 587-593), then one Skeleton object per label."""
  vox, rad, lens, gids, priv = bundle
  an = tuple(float(a) for a in anisotropy)
  t0 = time.perf_counter()
  seg_off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
  results = engine.assemble(vox, rad, seg_off, np.arange(lens.size, dtype=np.int64), shape, an, group_ids=gids,
                            seg_private=priv)
  if tm is not None:
    if vox.device.type == "cuda":
      torch.cuda.synchronize()
    tm["assemble"] = tm.get("assemble", 0.0) + time.perf_counter() - t0
  t0 = time.perf_counter()
  transform = np.array([[an[0], 0, 0, 0], [0, an[1], 0, 0], [0, 0, an[2], 0]], dtype=np.float32)
  out = {}
  for orig, (verts, edges, radii) in results.items():
    if verts.shape[0] == 0 or edges.shape[0] == 0:
      continue
    out[orig] = Skeleton._from_arrays(verts, edges, radii, orig, transform, "physical")
  if tm is not None:
    tm["finalize"] = tm.get("finalize", 0.0) + time.perf_counter() - t0
  return out


def connect_points(labels, start, end, anisotropy=(1, 1, 1), fill_holes=False, in_place=False, pdrf_scale=100000,
                   pdrf_exponent=4):
  """kimimaro.connect_points (kimimaro/intake.py:268-313) -> trace.point_to_point (kimimaro/trace.py:358-390): the
  centreline between two chosen voxels of a binary image.  EDT with a black border, DAF from `start`, PDRF, then
  dijkstra3d.dijkstra(PDRF, end, start): here the node-weighted distance field from `end` (b2t_edf_labels with the PDRF
  as node weights) and the walk from `start` along parents by rule T3 (the fix_branching=False machinery of the path
  kernel, one path).  The Skeleton comes back like Skeleton.from_path: vertices in PATH order (end first), edges
  (i, i+1), radii = DBF, physical units.  fill_holes is accepted and unused, like in the reference."""
  an = tuple(float(a) for a in anisotropy)
  start = tuple(int(v) for v in start)
  end = tuple(int(v) for v in end)
  lab = format_labels(np.asarray(labels).astype(bool), in_place=True)
  shape = lab.shape
  sx, sy, sz = shape
  start, end = start + (0,) * (3 - len(start)), end + (0,) * (3 - len(end))
  lin = lambda p: int(p[0]) + sx * (int(p[1]) + sy * int(p[2]))
  d_labels = _upload(lab.reshape(-1, order="F"))
  d_cc, n_cc = engine.connected_components(d_labels, shape)
  c_start, c_end = int(d_cc[lin(start)].item()), int(d_cc[lin(end)].item())
  if c_start == 0 or c_start != c_end:
    raise ValueError("Cannot extract centerline from disconnected components.")
  d_dbf = edt(d_labels, shape, an, black_border=True)
  dbf_max = np.float32(d_dbf.max().item())                    # of the whole image, like np.max(DBF) in point_to_point
  n_fg = int((d_cc == c_start).sum().item())
  jobs = engine.Jobs([c_start], [n_fg], [lin(start)], [lin(end)], [dbf_max], tb={0: [lin(start)]}, daf_source=[lin(start)])
  jobs.single_path = True
  params = dict(engine.TRACE_DEFAULTS)
  params.update(pdrf_scale=pdrf_scale, pdrf_exponent=pdrf_exponent, fix_branching=False)
  vox, rad, seg_off, _, _ = engine.trace_arena(d_cc, d_dbf, shape, an, jobs, params, n_cc)
  h_vox = vox.cpu().numpy().view(np.uint32).astype(np.int64)
  keep = h_vox != 0xFFFFFFFF
  path = h_vox[keep]
  radii = rad.cpu().numpy()[keep]
  z, r = np.divmod(path, sx * sy)
  y, x = np.divmod(r, sx)
  skel = Skeleton.from_path(np.stack([x, y, z], axis=1))      # trace.py:386: vertices in path order, edges (i, i+1)
  skel.radii = radii.astype(np.float32)                       # trace.py:388-389: DBF at the vertices
  skel.vertices *= np.array(an, dtype=np.float32)             # intake.py:310-311
  skel.space = "physical"
  return skel


@functools.wraps(_skeletonize)
def skeletonize(*args, **kwargs):
  # A full (generation 2) pass of CPython's cyclic collector walks every tracked object of the process
  # (torch alone brings hundreds of thousands) and costs 30-40 ms -- a quarter of a 512^3 pass -- so it is
  # suspended for the duration of the call and restored afterwards.  Results are unaffected.
  was_enabled = gc.isenabled()
  gc.disable()
  try:
    return _skeletonize(*args, **kwargs)
  finally:
    if was_enabled:
      gc.enable()
