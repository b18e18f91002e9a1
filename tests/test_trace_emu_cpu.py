"""The engine's invalidation (kimimaro_b200/csrc/trace.cu: invalidate_window, invalidate_strict, invalidate) on the CPU:
the device functions are compiled by g++ against a SIMT emulation (tests/host/emu_include/cuda_runtime.h: one OS thread
per CUDA thread of a 512-thread block, barriers for __syncthreads and the warp intrinsics) and must reproduce the
oracle's restatement of the same claim order voxel for voxel -- the key-ordered rounds the engine runs by default
(orc_invalidate_window), the literal heap of the strict mode (orc_invalidate_heap, which is the reference's compiled
extension voxel for voxel: test_oracle_cpu.py::test_invalidation_vs_reference_ext) and the hop-synchronous rounds
(orc_invalidate_rounds)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
from tests.test_oracle_cpu import _tube

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "host", "trace_emu.cpp")
DEPS = [SRC, os.path.join(HERE, "host", "emu_include", "cuda_runtime.h"),
        os.path.join(ROOT, "kimimaro_b200", "csrc", "trace.cu"), os.path.join(ROOT, "kimimaro_b200", "csrc", "common.cuh")]
OUT = os.path.join(ROOT, "oracle", "_cache", "trace_emu.so")


@pytest.fixture(scope="module")
def emu():
  os.makedirs(os.path.dirname(OUT), exist_ok=True)
  if (not os.path.exists(OUT)) or os.path.getmtime(OUT) < max(os.path.getmtime(d) for d in DEPS):
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wno-attributes",
                           "-I" + os.path.join(HERE, "host", "emu_include"), SRC, "-o", OUT])
  lib = ctypes.CDLL(OUT)
  lib.emu_invalidate.restype = ctypes.c_long
  return lib


ROUNDS, WINDOW, STRICT = 0, 1, 2


@pytest.fixture(scope="module")
def emu_tight():
  """The same harness with railroad's pair lists shrunk 32-fold (B2T_RR_CAP_SHIFT): lists overflow and the pile is rebuilt
  from the touched voxels again and again."""
  out = OUT.replace("trace_emu.so", "trace_emu_tight.so")
  if (not os.path.exists(out)) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in DEPS):
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wno-attributes",
                           "-DB2T_RR_CAP_SHIFT=5", "-I" + os.path.join(HERE, "host", "emu_include"), SRC, "-o", out])
  lib = ctypes.CDLL(out)
  lib.emu_invalidate.restype = ctypes.c_long
  return lib


def _engine(lib, vol, dbf, path, scale, const, an, window, mode=None, spill_words=0, team=0):
  sx, sy, sz = vol.shape
  cc = np.ascontiguousarray(vol.reshape(-1, order="F").astype(np.uint32))
  d = np.ascontiguousarray(dbf.reshape(-1, order="F").astype(np.float32))
  claim = np.full(cc.size, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
  p = np.asarray(path, dtype=np.int64).reshape(-1, 3)
  seeds = np.ascontiguousarray((p[:, 0] + sx * (p[:, 1] + sy * p[:, 2])).astype(np.uint32))
  delta = np.float32(window * min(an)) if window else np.float32(0)
  if mode is None:
    mode = WINDOW if window else ROUNDS
  n = lib.emu_invalidate(oracle._p(cc), oracle._p(d), oracle._p(claim), sx, sy, sz, ctypes.c_float(an[0]),
                         ctypes.c_float(an[1]), ctypes.c_float(an[2]), 1, int(vol.sum()), oracle._p(seeds), int(seeds.size),
                         ctypes.c_float(scale), ctypes.c_float(const), ctypes.c_float(delta), mode,
                         ctypes.c_long(spill_words), team)
  mask = ((claim != 0) & (cc == 1)).astype(np.uint8).reshape(vol.shape, order="F")
  return int(n), mask


@pytest.mark.parametrize("team", [0, 1])
@pytest.mark.parametrize("window", [0, 1.0, 0.5])
def test_engine_invalidation_equals_oracle(emu, window, team):
  rng = np.random.default_rng(17)
  mode = f"window:{window:g}" if window else "rounds"
  for trial in range(10):
    vol, path = _tube(rng)
    an = (16.0, 16.0, 40.0) if trial % 2 else (1.0, 1.0, 1.0)
    dbf = oracle.edt(vol, an, False)
    scale, const = float(rng.choice([1.0, 1.5, 4.0])), float(rng.choice([0, 1, 3])) * an[0]
    if trial == 3:                                  # seeds that are already invalid never expand (hpp:297-299)
      hole = path[len(path) // 2]
      vol = vol.copy(order="F")
      vol[hole] = 0
    ref = vol.copy(order="F")
    n_ref, ref = oracle.roll_invalidation_ball_inside_component(ref, dbf, scale, const, an, path, mode=mode)
    n, mask = _engine(emu, vol, dbf, path, scale, const, an, window, team=team)
    assert n == n_ref, (trial, n, n_ref)
    assert np.array_equal(mask, ref), (trial, int((mask != ref).sum()))


def test_engine_strict_invalidation_equals_the_literal_heap(emu, ref_ext):
  """invalidate_strict (one warp emulating std::priority_queue with libstdc++'s push_heap / pop_heap and the reference's
  `>=` comparator) against the oracle's literal form AND, when it is built, against the reference's own compiled
  extension: the same voxels, not merely the same count -- on shapes where tie order decides (isotropic grids, seeds at
  equal distances, volumes whose x extent puts path voxels on the x faces where the corner entries alias)."""
  rng = np.random.default_rng(23)
  for trial in range(8):
    vol, path = _tube(rng)
    an = (16.0, 16.0, 40.0) if trial % 2 else (1.0, 1.0, 1.0)
    if trial >= 6:                                  # crop so that the tube touches the x faces
      xs = np.flatnonzero(vol.any(axis=(1, 2)))
      lo, hi = xs[0] + 2, xs[-1] - 1
      vol = np.asfortranarray(vol[lo:hi])
      path = [(p[0] - lo, p[1], p[2]) for p in path if lo <= p[0] < hi]
    dbf = oracle.edt(vol, an, False)
    scale, const = float(rng.choice([1.0, 1.5, 4.0])), float(rng.choice([0, 1, 3])) * an[0]
    ref = vol.copy(order="F")
    n_ref, ref = oracle.roll_invalidation_ball_inside_component(ref, dbf, scale, const, an, path, mode="heap")
    n, mask = _engine(emu, vol, dbf, path, scale, const, an, 0, mode=STRICT, spill_words=3 * (28 * int(vol.sum()) + 128))
    assert n == n_ref, (trial, n, n_ref)
    assert np.array_equal(mask, ref), (trial, int((mask != ref).sum()))
    if ref_ext is not None:
      ext = np.asfortranarray(vol.astype(bool))
      n_ext, ext = ref_ext.roll_invalidation_ball_inside_component(ext, dbf, scale, const, np.asarray(an, np.float32),
                                                                   [tuple(int(c) for c in p) for p in path])
      assert n == int(n_ext) and np.array_equal(mask, ext.view(np.uint8)), trial


def test_engine_strict_heap_spills_and_reports_capacity(emu):
  """A fat label whose ball covers it entirely: the heap outgrows the label's own region (4 entries per voxel), moves to
  the spill arena and the result is still the literal heap's; without a spill arena the call reports B2T_ERR_CAPACITY."""
  vol = np.zeros((26, 26, 26), np.uint8, order="F")
  vol[1:25, 1:25, 1:25] = 1
  an = (1.0, 1.0, 1.0)
  dbf = oracle.edt(vol, an, False)
  path = [(12, 12, z) for z in range(4, 21)]
  ref = vol.copy(order="F")
  n_ref, ref = oracle.roll_invalidation_ball_inside_component(ref, dbf, 4.0, 10.0, an, path, mode="heap")
  stats = (ctypes.c_int64 * 2).in_dll(oracle.lib(), "orc_heap_stats")
  assert stats[0] > int(vol.sum()) + 64, "the case must outgrow the harness's static region (1 per voxel + 64)"
  n, mask = _engine(emu, vol, dbf, path, 4.0, 10.0, an, 0, mode=STRICT, spill_words=3 * (28 * int(vol.sum()) + 128))
  assert n == n_ref and np.array_equal(mask, ref)
  n, _ = _engine(emu, vol, dbf, path, 4.0, 10.0, an, 0, mode=STRICT, spill_words=0)
  assert n == -4


# ---- the whole path loop (trace_kernel) on the emulated block against oracle.teasar.trace ----
def _emulated_paths(lib, cc, n_cc, all_dbf, an, params, window, fix_borders_targets=None, mode=None, n_team=0):
  """Mirrors kimimaro_b200/engine.py:trace_arena_start for host arrays: per-label root / DAF / PDRF from the oracle's
  pieces (the GPU gets them from field.cu, verified on the GPU), one DAF bucket per label, then trace_kernel."""
  from oracle import teasar
  from kimimaro_b200.engine import DESC_DTYPE
  sx, sy, sz = cc.shape
  V = cc.size
  flat = lambda a, dt: np.ascontiguousarray(a.reshape(-1, order="F").astype(dt))
  pdrf = np.zeros(V, np.float32)
  keys_l, hist, cursor, descs, targets = [], [], [], [], []
  lin = lambda p: int(p[0]) + sx * (int(p[1]) + sy * int(p[2]))
  region = path_off = 0
  for l in range(1, n_cc + 1):
    labels = np.asfortranarray(cc == l)
    n_fg = int(labels.sum())
    DBF = np.where(labels, all_dbf, 0).astype(np.float32, order="F")
    dbf_max = np.max(DBF)
    root = teasar.find_root(labels.view(np.uint8), an)
    DBFi = DBF.copy(order="F")
    DBFi[DBFi == 0] = np.inf
    DAF, target = oracle.euclidean_distance_field(labels.view(np.uint8), root, anisotropy=an, free_space_radius=0,
                                                  return_max_location=True)
    DAF[DAF == np.inf] = 0
    P = teasar.compute_pdrf(dbf_max, params["pdrf_scale"], params["pdrf_exponent"], DBFi, DAF, DAF[target])
    idx = np.flatnonzero(labels.reshape(-1, order="F"))
    pdrf[idx] = P.reshape(-1, order="F")[idx]
    daf_bits = DAF.reshape(-1, order="F")[idx].astype(np.float32).view(np.uint32).astype(np.uint64)
    keys_l.append((daf_bits << np.uint64(32)) | idx.astype(np.uint64))
    d = np.zeros(1, DESC_DTYPE)
    cap = 2 * n_fg + 2 + 64
    d["segid"], d["root"], d["n_fg"], d["region_off"], d["path_off"], d["path_cap"] = l, lin(root), n_fg, region, path_off, cap
    d["tb_off"], d["tb_n"], d["ta_off"], d["ta_n"] = len(targets), 1, len(targets) + 1, 0
    d["max_paths"], d["bucket_row"] = 0xFFFFFFFF, l - 1
    d["bbox_x0"], d["bbox_x1"] = 0, sx - 1       # the array the oracle's trace() below runs on is the whole volume
    targets.append(lin(target))
    descs.append(d)
    hist.append(n_fg)
    region += n_fg
    path_off += cap
    cursor.append(region)                      # END of the label's only bucket inside keys
  desc = np.concatenate(descs)
  keys = np.ascontiguousarray(np.concatenate(keys_l))
  hist = np.array(hist + [0], np.uint32)
  cursor = np.array(cursor + [0], np.uint32)
  targets = np.array(targets + [0], np.uint32)
  ccf, dbf = flat(cc, np.uint32), flat(all_dbf, np.float32)
  dist = np.full(V, np.inf, np.float32)
  claim = np.full(V, 0xFFFFFFFFFFFFFFFF, np.uint64)
  stamp = np.zeros(V, np.uint32)
  scratch = np.zeros(22 * region + 16, np.uint32)
  paths = np.zeros(path_off + 8, np.uint32)
  n = len(desc)
  out_len, out_np, out_status = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.int32)
  out_stats, counter = np.zeros(4 * n, np.uint32), np.zeros(1, np.uint32)
  p, cf = oracle._p, ctypes.c_float
  if mode is None:
    mode = WINDOW if window else ROUNDS
  lib.b2t_trace_heap_words.restype = ctypes.c_uint64
  heap_static = int(lib.b2t_trace_heap_words(ctypes.c_uint64(region), ctypes.c_uint64(n))) if mode == STRICT else 0
  heap = np.zeros(heap_static + 3 * 29 * int(sum(hist)) + 256, np.uint32) if mode == STRICT else np.zeros(1, np.uint32)
  lib.emu_trace_batch(p(ccf), p(dbf), p(pdrf), p(dist), p(claim), p(stamp), sx, sy, sz, cf(an[0]), cf(an[1]), cf(an[2]),
                      p(desc), n, cf(params["scale"]), cf(params["const"]), cf(0.5), cf(0.0), 1, 1, p(keys), p(hist),
                      p(cursor), p(scratch), p(paths), p(targets), p(out_len), p(out_np), p(out_status), p(out_stats),
                      p(counter), mode, cf(window * min(an)), p(heap), ctypes.c_uint64(heap.size if mode == STRICT else 0),
                      ctypes.c_uint64(heap_static), n_team)
  assert (out_status == 0).all(), out_status
  got = {}
  for i in range(n):
    seg = paths[int(desc[i]["path_off"]): int(desc[i]["path_off"]) + int(out_len[i])]
    cuts = np.flatnonzero(seg == 0xFFFFFFFF)
    got[i + 1] = [a for a in np.split(seg, cuts + 1)[:-1]]
    got[i + 1] = [a[:-1].astype(np.int64) for a in got[i + 1]]
    assert len(got[i + 1]) == int(out_np[i])
  return got


@pytest.mark.parametrize("n_team", [0, 2])
@pytest.mark.parametrize("window", [0, 1.0, "strict"])
def test_engine_path_loop_equals_oracle(emu, window, n_team):
  """find_target -> railroad -> invalidate -> rail, label after label, on the source text the GPU runs: every path of
  every label must be the oracle's, voxel for voxel and in the same order, under the key-ordered claim the engine ships
  (window 1, oracle mode 'window:1'), under the strict mode (oracle mode 'heap' == the compiled reference) and under the
  hop-synchronous rounds (window 0)."""
  from oracle import teasar
  from tests.synth import synthetic_tubes
  emode = None
  if window == "strict":
    mode, window, emode = "heap", 0, STRICT
  else:
    mode = f"window:{window:g}" if window else "rounds"
  params = dict(teasar.DEFAULT_TEASAR_PARAMS)
  params.update(scale=1.5, const=30)                       # small tubes: several paths per label
  n_paths = 0
  for seed, shape, n_tubes, an in ((5, (48, 40, 32), 5, (16.0, 16.0, 40.0)), (6, (40, 40, 40), 4, (1.0, 1.0, 1.0))):
    lab = synthetic_tubes(shape, n_tubes, seed=seed)
    cc, n_cc = oracle.connected_components(lab)
    keep = [l for l in range(1, n_cc + 1) if (cc == l).sum() > 200]
    cc = np.asfortranarray(np.where(np.isin(cc, keep), cc, 0))
    cc, n_cc = oracle.connected_components(cc)
    all_dbf = oracle.edt(cc, an, False)
    got = _emulated_paths(emu, cc, n_cc, all_dbf, an, params, window, mode=emode, n_team=min(n_team, n_cc))
    sx, sy, sz = cc.shape
    for l in range(1, n_cc + 1):
      labels = np.asfortranarray(cc == l)
      DBF = np.where(labels, all_dbf, 0).astype(np.float32, order="F")
      _, ref = teasar.trace(labels, DBF, anisotropy=an, invalidation_mode=mode, return_paths=True, **params)
      ref = [(np.asarray(q, np.int64)[:, 0] + sx * (np.asarray(q, np.int64)[:, 1] + sy * np.asarray(q, np.int64)[:, 2]))
             for q in ref if len(q) > 0]
      assert len(got[l]) == len(ref), (seed, l, len(got[l]), len(ref))
      for a, b in zip(got[l], ref):
        assert np.array_equal(a, b), (seed, l)
      n_paths += len(ref)
  assert n_paths >= 10


def test_railroad_pile_rebuild_after_overflow(emu_tight):
  """railroad keeps superseded pairs in its lists (lazy deletion); when a list fills up the pile is rebuilt from the touched
  voxels.  With the lists shrunk 32-fold that happens many times per search, and the paths still have to be the oracle's."""
  from oracle import teasar
  from tests.synth import synthetic_tubes
  params = dict(teasar.DEFAULT_TEASAR_PARAMS)
  params.update(scale=1.5, const=30)
  an = (16.0, 16.0, 40.0)
  lab = synthetic_tubes((64, 56, 40), 4, seed=8)
  cc, n_cc = oracle.connected_components(lab)
  keep = [l for l in range(1, n_cc + 1) if (cc == l).sum() > 1500]
  cc = np.asfortranarray(np.where(np.isin(cc, keep), cc, 0))
  cc, n_cc = oracle.connected_components(cc)
  assert n_cc >= 1
  all_dbf = oracle.edt(cc, an, False)
  before = ctypes.c_int.in_dll(emu_tight, "g_emu_rr_rebuilds").value
  got = _emulated_paths(emu_tight, cc, n_cc, all_dbf, an, params, 1.0)
  assert ctypes.c_int.in_dll(emu_tight, "g_emu_rr_rebuilds").value > before, "the case must overflow to test anything"
  sx, sy, sz = cc.shape
  for l in range(1, n_cc + 1):
    labels = np.asfortranarray(cc == l)
    DBF = np.where(labels, all_dbf, 0).astype(np.float32, order="F")
    _, ref = teasar.trace(labels, DBF, anisotropy=an, invalidation_mode="window:1", return_paths=True, **params)
    ref = [(np.asarray(q, np.int64)[:, 0] + sx * (np.asarray(q, np.int64)[:, 1] + sy * np.asarray(q, np.int64)[:, 2]))
           for q in ref if len(q) > 0]
    assert len(got[l]) == len(ref)
    for a, b in zip(got[l], ref):
      assert np.array_equal(a, b), l
