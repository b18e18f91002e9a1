"""
TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

CPU oracle for the B200 TEASAR engine: ctypes wrappers over oracle/oracle.c (a C restatement
of the reference's hot-path arithmetic) plus a numpy conductor that mirrors
kimimaro/trace.py and kimimaro/intake.py (see oracle/teasar.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may
import this package.  kimimaro_b200 (the product) never does.

Parity status:
  * invalidation / target finder / border targets: PINNED against the reference's own in-tree
    extension compiled from /root/reference into oracle/_ref (oracle/build_ref.py).
  * EDT / Dijkstra family / fill / CCL: restated from the published algorithms of the
    un-vendored PyPI deps (edt>=3.0.0, dijkstra3d>=1.15.0, fill-voids>=2.0.0,
    connected-components-3d>=3.16.0, requirements.txt:2-7); pinned only through brute-force
    definitions and the reference's known-answer tests (automated_test.py:48-199).
    "parity unpinned vs the binaries".
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(HERE, "liboracle.so")
_lib = None

c_i64 = ctypes.c_int64
c_f32 = ctypes.c_float
c_int = ctypes.c_int
c_void_p = ctypes.c_void_p


def build(force=False):
  src = os.path.join(HERE, "oracle.c")
  if (not force) and os.path.exists(_LIB_PATH) and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src):
    return _LIB_PATH
  # -ffp-contract=off: float32 expressions must round exactly like the reference's scalar C++
  cmd = ["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fvisibility=hidden", "-shared", "-fPIC",
         src, "-o", _LIB_PATH, "-lm"]
  subprocess.check_call(cmd)
  return _LIB_PATH


def lib():
  global _lib
  if _lib is None:
    build()
    _lib = ctypes.CDLL(_LIB_PATH)
    _lib.orc_edf.restype = c_i64
    _lib.orc_railroad.restype = c_i64
    _lib.orc_invalidate_seq.restype = c_i64
    _lib.orc_invalidate_rounds.restype = c_i64
    _lib.orc_invalidate_heap.restype = c_i64
    _lib.orc_invalidate_window.restype = c_i64
    _lib.orc_fill_voids.restype = c_i64
    _lib.orc_ccl26.restype = c_i64
  return _lib


def _p(a):
  return a.ctypes.data_as(c_void_p)


def _shape3(a):
  s = tuple(a.shape) + (1,) * (3 - a.ndim)
  return s


def _f(a, dtype):
  """Fortran-ordered, dtype-converted, 3-D view/copy."""
  a = np.asarray(a)
  while a.ndim < 3:
    a = a[..., np.newaxis]
  return np.asfortranarray(a, dtype=dtype)


# --------------------------------------------------------------------------------------------
def edt(labels, anisotropy=(1, 1, 1), black_border=False, brute_force=False):
  """edt.edt(labels, anisotropy, black_border) -> float32, Fortran order.  2-D input runs the
  2-D transform (x,y passes) like the library does (intake.py:565)."""
  labels = np.asarray(labels)
  ndim = labels.ndim
  assert ndim in (2, 3)
  an = tuple(float(a) for a in anisotropy) + (1.0,) * (3 - len(anisotropy))
  if labels.dtype == bool:
    labels = labels.view(np.uint8)
  if labels.dtype.itemsize > 4:
    _, inv = np.unique(labels, return_inverse=True)  # order-preserving: 0 stays the minimum
    lab32 = inv.reshape(labels.shape).astype(np.uint32)
    if labels.min() != 0:
      lab32 += 1
    labels = lab32
  L = _f(labels, np.uint32)
  sx, sy, sz = L.shape
  out = np.zeros(L.shape, dtype=np.float32, order="F")
  fn = lib().orc_edt_bruteforce if brute_force else lib().orc_edt
  fn(_p(L), c_i64(sx), c_i64(sy), c_i64(sz), c_f32(an[0]), c_f32(an[1]), c_f32(an[2]),
     c_int(int(bool(black_border))), c_int(ndim), _p(out))
  return out.reshape(labels.shape, order="F") if ndim == 2 else out


def euclidean_distance_field(field, source, anisotropy=(1, 1, 1), free_space_radius=0.0,
                             return_max_location=False):
  """dijkstra3d.euclidean_distance_field (trace.py:139-145, 302-307)."""
  F = _f(np.asarray(field) != 0, np.uint8)
  sx, sy, sz = F.shape
  src = np.atleast_2d(np.asarray(source, dtype=np.int64))
  lin = (src[:, 0] + sx * (src[:, 1] + sy * src[:, 2])).astype(np.int64)
  dist = np.empty(F.shape, dtype=np.float32, order="F")
  mx = lib().orc_edf(_p(F), c_i64(sx), c_i64(sy), c_i64(sz),
                     c_f32(anisotropy[0]), c_f32(anisotropy[1]), c_f32(anisotropy[2]),
                     _p(lin), c_i64(lin.size), c_f32(float(free_space_radius)), _p(dist))
  if return_max_location:
    return dist, tuple(int(v) for v in np.unravel_index(mx, F.shape, order="F"))
  return dist


def railroad(field, source):
  """dijkstra3d.railroad(field, source) (trace.py:240-242): path[0] = rail voxel, path[-1] = source."""
  F = _f(field, np.float32)
  sx, sy, sz = F.shape
  s = int(source[0]) + sx * (int(source[1]) + sy * int(source[2]))
  dist = np.empty(F.size, dtype=np.float32)
  path = np.empty(F.size + 2, dtype=np.int64)
  n = lib().orc_railroad(_p(F), c_i64(sx), c_i64(sy), c_i64(sz), c_i64(s), _p(dist), _p(path))
  idx = path[:n]
  return np.stack(np.unravel_index(idx, F.shape, order="F"), axis=1).astype(np.uint32)


def parental_field(field, source):
  """dijkstra3d.parental_field (trace.py:155)."""
  F = _f(field, np.float32)
  sx, sy, sz = F.shape
  s = int(source[0]) + sx * (int(source[1]) + sy * int(source[2]))
  dist = np.empty(F.size, dtype=np.float32)
  parents = np.zeros(F.shape, dtype=np.uint32, order="F")
  lib().orc_parental_field(_p(F), c_i64(sx), c_i64(sy), c_i64(sz), c_i64(s), _p(dist), _p(parents))
  return parents


def path_from_parents(parents, target):
  """dijkstra3d.path_from_parents (trace.py:244): source -> target order."""
  P = parents.ravel(order="F")
  sx, sy, sz = parents.shape
  loc = int(target[0]) + sx * (int(target[1]) + sy * int(target[2]))
  out = []
  while P[loc]:
    out.append(loc)
    loc = int(P[loc]) - 1
  out.append(loc)
  idx = np.array(out[::-1], dtype=np.int64)
  return np.stack(np.unravel_index(idx, parents.shape, order="F"), axis=1).astype(np.uint32)


def invalidation_radii(DBF, scale, const, path):
  """max_distances of skeletontricks.pyx:393-395 under NumPy-2 scalar rules:
  fl32(fl32(scale*DBF[v]) + const) (verified on the compiled extension)."""
  path = np.asarray(path, dtype=np.int64).reshape(-1, 3)
  d = DBF[path[:, 0], path[:, 1], path[:, 2]].astype(np.float32)
  return (np.float32(scale) * d + np.float32(const)).astype(np.float32)


WINDOW_ROUNDS = [0, 0]   # (rounds, calls) of the 'window:' mode, for the design study


def roll_invalidation_ball_inside_component(labels, DBF, scale, const, anisotropy, path, mode="rounds"):
  """skeletontricks.pyx:373-418.  labels (uint8/bool, Fortran) is edited in place.
  mode: 'rounds' = the engine's round-synchronous claim; 'seq' = ordered best-first claim with canonical ties;
  'heap' = the reference's loop literally, libstdc++'s heap order for equal keys included (== the compiled reference)."""
  assert labels.flags["F_CONTIGUOUS"] and labels.ndim == 3
  m = labels.view(np.uint8)
  sx, sy, sz = m.shape
  path = np.asarray(path, dtype=np.int64).reshape(-1, 3)
  radii = invalidation_radii(DBF, scale, const, path)
  seeds = (path[:, 0] + sx * (path[:, 1] + sy * path[:, 2])).astype(np.int64)
  if isinstance(mode, str) and mode.startswith("window:"):     # 'window:<delta in units of the smallest voxel edge>'
    parts = mode.split(":")                                    # 'window:<delta>[:hi]' (hi: the later seed wins a tie)
    delta = np.float32(float(parts[1]) * float(min(anisotropy)))
    nr = c_i64(0)
    n = lib().orc_invalidate_window(_p(m), c_i64(sx), c_i64(sy), c_i64(sz), c_f32(anisotropy[0]), c_f32(anisotropy[1]),
                                    c_f32(anisotropy[2]), _p(seeds), _p(radii), c_i64(seeds.size), c_f32(delta),
                                    ctypes.byref(nr), ctypes.c_int(1 if len(parts) > 2 and parts[2] == "hi" else 0))
    WINDOW_ROUNDS[0] += nr.value
    WINDOW_ROUNDS[1] += 1
    return int(n), labels
  fn = {"rounds": lib().orc_invalidate_rounds, "seq": lib().orc_invalidate_seq, "heap": lib().orc_invalidate_heap}[mode]
  n = fn(_p(m), c_i64(sx), c_i64(sy), c_i64(sz), c_f32(anisotropy[0]), c_f32(anisotropy[1]),
         c_f32(anisotropy[2]), _p(seeds), _p(radii), c_i64(seeds.size))
  return int(n), labels


def fill_voids(mask):
  """fill_voids.fill(mask, in_place=True, return_fill_count=True) (trace.py:109)."""
  assert mask.flags["F_CONTIGUOUS"] and mask.ndim == 3
  m = mask.view(np.uint8)
  n = lib().orc_fill_voids(_p(m), c_i64(m.shape[0]), c_i64(m.shape[1]), c_i64(m.shape[2]))
  return mask, int(n)


def connected_components(labels):
  """cc3d.connected_components(labels) 26-connected multi-label (utility.py:77); 2-D -> 8-connected."""
  labels = np.asarray(labels)
  shape = labels.shape
  if labels.dtype == bool:
    labels = labels.view(np.uint8)
  if labels.dtype.itemsize > 4 or labels.dtype.kind == "i":
    _, inv = np.unique(labels, return_inverse=True)
    lab32 = inv.reshape(labels.shape).astype(np.uint32)
    if labels.min() != 0:
      lab32 += 1
    labels = lab32
  L = _f(labels, np.uint32)
  out = np.zeros(L.shape, dtype=np.uint32, order="F")
  n = lib().orc_ccl26(_p(L), c_i64(L.shape[0]), c_i64(L.shape[1]), c_i64(L.shape[2]), _p(out))
  return out.reshape(shape, order="F"), int(n)
