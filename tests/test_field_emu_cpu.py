"""kimimaro_b200/csrc/field.cu on the CPU: label statistics, the multi-source distance-field sweep (cooperative kernel,
emulated as one block of 1024 threads), the per-label arg-max, PDRF + target buckets and the grid-wide ball invalidation
are compiled by g++ against the SIMT emulation and run through the library's own entry points on host arrays against the
oracle (binary-heap Dijkstra, numpy PDRF, heap-ordered invalidation)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "host", "field_emu.cpp")
DEPS = [SRC, os.path.join(HERE, "host", "emu_include", "cuda_runtime.h"), os.path.join(HERE, "host", "emu_include", "simt_impl.h"),
        os.path.join(HERE, "host", "emu_include", "cooperative_groups.h"),
        os.path.join(ROOT, "kimimaro_b200", "csrc", "field.cu"), os.path.join(ROOT, "kimimaro_b200", "csrc", "common.cuh")]
OUT = os.path.join(ROOT, "oracle", "_cache", "field_emu.so")
c_i64, c_u64, c_u32, c_f32 = ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_float
p = oracle._p


def _build(out, extra=()):
  os.makedirs(os.path.dirname(out), exist_ok=True)
  if (not os.path.exists(out)) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in DEPS):
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wno-attributes",
                           *extra, "-I" + os.path.join(HERE, "host", "emu_include"), "-I" + os.path.join(HERE, "host"), SRC,
                           "-o", out])
  return ctypes.CDLL(out)


@pytest.fixture(scope="module")
def emu():
  return _build(OUT)


@pytest.fixture(scope="module")
def emu_tiny_lists():
  """the same library with the solo sweep's shared-memory frontier shrunk to 8 pairs: almost every round spills into the
  label's global queue (the stamp-deduplicated path)"""
  return _build(OUT.replace(".so", "_cap8.so"), extra=("-DB2T_EDF_SOLO=1", "-DB2T_EDF_SOLO_CAP=8"))


def _volume(seed, shape=(40, 36, 28), n=5):
  from tests.synth import synthetic_tubes
  lab = synthetic_tubes(shape, n, seed=seed)
  cc, n_cc = oracle.connected_components(lab)
  return np.asfortranarray(cc), n_cc


def test_label_stats(emu):
  cc, n = _volume(31)
  an = (16.0, 16.0, 40.0)
  dbf = oracle.edt(cc, an, False)
  sx, sy, sz = cc.shape
  ccf = np.ascontiguousarray(cc.reshape(-1, order="F").astype(np.uint32))
  dbff = np.ascontiguousarray(dbf.reshape(-1, order="F"))
  count, bbox = np.empty(n + 1, np.uint32), np.empty(6 * (n + 1), np.int32)
  dbfmax, first = np.empty(n + 1, np.float32), np.empty(n + 1, np.uint32)
  assert emu.b2t_label_stats(p(ccf), p(dbff), c_i64(sx), c_i64(sy), c_i64(sz), c_u32(n), p(count), p(bbox), p(dbfmax),
                             p(first), None) == 0
  bbox = bbox.reshape(-1, 6)
  for l in range(1, n + 1):
    m = cc == l
    idx = np.argwhere(m)
    assert count[l] == m.sum()
    assert list(bbox[l]) == list(idx.min(axis=0)) + list(idx.max(axis=0))
    assert dbfmax[l] == dbf[m].max()
    assert first[l] == np.flatnonzero(m.reshape(-1, order="F"))[0]      # first_label (pyx:307-326)


def _edf(emu, ccf, shape, an, sources, node_w=None):
  sx, sy, sz = shape
  V = ccf.size
  dist = np.full(V, np.inf, np.float32)
  stamp = np.zeros(V, np.uint32)
  cap = int((ccf != 0).sum()) + 8
  queue = np.zeros(2 * cap, np.uint32)
  ctrl = np.zeros(16, np.uint32)
  src = np.ascontiguousarray(np.asarray(sources, np.uint32))
  rc = emu.b2t_edf_multi(p(ccf), c_i64(sx), c_i64(sy), c_i64(sz), c_f32(an[0]), c_f32(an[1]), c_f32(an[2]), p(src),
                         c_u32(src.size), c_f32(0.0), c_u32(0), p(node_w) if node_w is not None else None, p(dist), p(stamp),
                         p(queue), c_u64(cap), p(ctrl), None)
  assert rc == 0
  return dist


@pytest.mark.parametrize("an,nb", [((16.0, 16.0, 40.0), 16), ((1.0, 1.0, 1.0), 16), ((16.0, 16.0, 40.0), 1024)])
def test_distance_field_argmax_pdrf(emu, an, nb):
  """One sweep for ALL labels (each from its own source) must equal the oracle's Dijkstra field label by label, bit for
  bit; then the per-label arg-max (rule T2) and the fused PDRF against the oracle's numpy compute_pdrf."""
  from oracle import teasar
  cc, n = _volume(32)
  sx, sy, sz = cc.shape
  ccf = np.ascontiguousarray(cc.reshape(-1, order="F").astype(np.uint32))
  all_dbf = oracle.edt(cc, an, False)
  roots, ref_daf, ref_target = [], {}, {}
  for l in range(1, n + 1):
    labels = np.asfortranarray(cc == l).view(np.uint8)
    root = teasar.find_root(labels, an)
    daf, target = oracle.euclidean_distance_field(labels, root, anisotropy=an, free_space_radius=0, return_max_location=True)
    roots.append(int(root[0]) + sx * (int(root[1]) + sy * int(root[2])))
    ref_daf[l], ref_target[l] = daf, target
  dist = _edf(emu, ccf, cc.shape, an, roots)
  got = dist.reshape(cc.shape, order="F")
  for l in range(1, n + 1):
    m = cc == l
    assert np.array_equal(got[m], ref_daf[l][m]), l
  best = np.zeros(n + 1, np.uint64)
  assert emu.b2t_field_argmax(p(ccf), p(dist), c_i64(sx), c_i64(sy), c_i64(sz), c_u32(n), p(best), None) == 0
  vals = (best >> np.uint64(32)).astype(np.uint32).view(np.float32)
  idx = (np.uint64(0xFFFFFFFF) - (best & np.uint64(0xFFFFFFFF))).astype(np.int64)
  for l in range(1, n + 1):
    t = ref_target[l]
    assert idx[l] == int(t[0]) + sx * (int(t[1]) + sy * int(t[2])) and vals[l] == ref_daf[l][tuple(t)], l
  # PDRF + buckets (engine.py mirrors: M and 1/maxdaf come from the host's numpy)
  from kimimaro_b200.engine import compute_M_array
  dbfmax = np.array([0] + [all_dbf[cc == l].max() for l in range(1, n + 1)], np.float32)
  M = np.zeros(n + 1, np.float32)
  M[1:] = compute_M_array(dbfmax[1:])
  inv = np.zeros(n + 1, np.float32)
  with np.errstate(all="ignore"):                          # trace.py:352-354: only when max DAF is not 0
    inv[1:] = np.where(vals[1:] != 0, np.float32(1) / vals[1:], np.float32(0)).astype(np.float32)
  row = np.concatenate(([0xFFFFFFFF], np.arange(n))).astype(np.uint32)      # table row of label l: l - 1
  V = ccf.size               # nb = 1024: the (label x bucket) table spans two tiles of the multi-block scan
  pdrf, claim = np.zeros(V, np.float32), np.zeros(V, np.uint64)
  hist, cursor = np.zeros((n + 1) * nb + 1, np.uint32), np.zeros((n + 1) * nb + 1, np.uint32)
  keys = np.zeros(int((ccf != 0).sum()) + 1, np.uint64)
  dbff = np.ascontiguousarray(all_dbf.reshape(-1, order="F"))
  params = teasar.DEFAULT_TEASAR_PARAMS
  assert emu.b2t_pdrf_and_buckets(p(ccf), p(dbff), p(dist), p(pdrf), p(claim), c_i64(sx), c_i64(sy), c_i64(sz), c_u32(n), p(M),
                                  p(inv), p(row), c_u32(n), c_f32(params["pdrf_scale"]), c_f32(params["pdrf_exponent"]), nb, p(hist),
                                  p(cursor), p(keys), None) == 0
  P = pdrf.reshape(cc.shape, order="F")
  for l in range(1, n + 1):
    m = cc == l
    DBF = np.where(m, all_dbf, 0).astype(np.float32, order="F")
    DBF[DBF == 0] = np.inf
    daf = ref_daf[l].copy()
    daf[daf == np.inf] = 0
    ref = teasar.compute_pdrf(dbfmax[l], params["pdrf_scale"], params["pdrf_exponent"], DBF, daf, daf[tuple(ref_target[l])])
    assert np.array_equal(P[m], ref[m]), l
    # every voxel of the label sits in exactly one bucket, buckets ordered by DAF
    rows = slice((l - 1) * nb, l * nb)
    assert hist[rows].sum() == m.sum()
    ends = cursor[rows]
    ks = [keys[e - h:e] for e, h in zip(ends, hist[rows])]
    assert sorted(int(k & np.uint64(0xFFFFFFFF)) for b in ks for k in b) == sorted(np.flatnonzero(m.reshape(-1, order="F")).tolist())
    tops = [(b >> np.uint64(32)).max() for b in ks if b.size]
    lows = [(b >> np.uint64(32)).min() for b in ks if b.size]
    assert all(tops[i] <= lows[i + 1] for i in range(len(tops) - 1))
  assert (claim[ccf != 0] == np.uint64(0xFFFFFFFFFFFFFFFF)).all()
  assert np.isinf(dist[ccf != 0]).all()                     # the scatter hands the distance field back as +inf


@pytest.mark.parametrize("n_team,node_w", [(0, False), (2, False), (5, False), (1, True)])
def test_distance_field_label_by_label(emu, n_team, node_w):
  """b2t_edf_labels: every label sweeps its own rounds -- a CTA per label, or a team for the first n_team jobs (emulated
  clusters have one CTA: what is checked is the team code path with its counters in global memory) -- and must give
  exactly the field of the grid-wide sweep and of the oracle's Dijkstra; with node weights, dijkstra3d.parental_field."""
  from oracle import teasar
  an = (16.0, 16.0, 40.0)
  cc, n = _volume(34, shape=(44, 40, 30), n=6)
  sx, sy, sz = cc.shape
  ccf = np.ascontiguousarray(cc.reshape(-1, order="F").astype(np.uint32))
  V = ccf.size
  counts = np.bincount(ccf, minlength=n + 1)
  roots = {}
  for l in range(1, n + 1):
    r = teasar.find_root(np.asfortranarray(cc == l).view(np.uint8), an)
    roots[l] = int(r[0]) + sx * (int(r[1]) + sy * int(r[2]))
  order = sorted(range(1, n + 1), key=lambda l: -counts[l])
  tab = np.zeros((n, 4), np.uint32)
  off = 0
  for i, l in enumerate(order):
    tab[i] = (roots[l], l, counts[l], off)
    off += counts[l]
  rng = np.random.default_rng(5)
  w = np.ascontiguousarray((rng.random(V) * 10 + 0.5).astype(np.float32)) if node_w else None
  dist = np.full(V, np.inf, np.float32)
  stamp = np.zeros(V, np.uint32)
  queue = np.zeros(2 * off + 8, np.uint32)
  ctrl = np.zeros(4 * max(n_team, 1), np.uint32)
  n_team = min(n_team, n)
  assert emu.b2t_edf_labels(p(ccf), c_i64(sx), c_i64(sy), c_i64(sz), c_f32(an[0]), c_f32(an[1]), c_f32(an[2]), p(tab), c_u32(n),
                            c_u32(n_team), p(w) if node_w else None, p(dist), p(stamp), p(queue), p(ctrl), None) == 0
  ref = _edf(emu, ccf, cc.shape, an, [roots[l] for l in range(1, n + 1)], node_w=w)
  assert np.array_equal(dist, ref)
  if not node_w:
    got = dist.reshape(cc.shape, order="F")
    for l in range(1, n + 1):
      labels = np.asfortranarray(cc == l).view(np.uint8)
      daf = oracle.euclidean_distance_field(labels, np.unravel_index(roots[l], cc.shape, order="F"), anisotropy=an)
      assert np.array_equal(got[cc == l], daf[cc == l]), l


def test_distance_field_solo_spill(emu, emu_tiny_lists):
  """edf_label_solo with a frontier list of 8 pairs: pairs in shared memory and spilled voxels in the global queue mix
  in almost every round, and the field must still be the one the full-size build computes, bit for bit."""
  an = (16.0, 16.0, 40.0)
  cc, n = _volume(35, shape=(44, 40, 30), n=6)
  sx, sy, sz = cc.shape
  ccf = np.ascontiguousarray(cc.reshape(-1, order="F").astype(np.uint32))
  counts = np.bincount(ccf, minlength=n + 1)
  order = sorted(range(1, n + 1), key=lambda l: -counts[l])
  tab = np.zeros((n, 4), np.uint32)
  off = 0
  for i, l in enumerate(order):
    tab[i] = (int(np.flatnonzero(ccf == l)[0]), l, counts[l], off)
    off += counts[l]
  out = []
  for lib_ in (emu, emu_tiny_lists):
    dist = np.full(ccf.size, np.inf, np.float32)
    stamp = np.zeros(ccf.size, np.uint32)
    queue = np.zeros(2 * off + 8, np.uint32)
    ctrl = np.zeros(4, np.uint32)
    assert lib_.b2t_edf_labels(p(ccf), c_i64(sx), c_i64(sy), c_i64(sz), c_f32(an[0]), c_f32(an[1]), c_f32(an[2]), p(tab),
                               c_u32(n), c_u32(0), None, p(dist), p(stamp), p(queue), p(ctrl), None) == 0
    out.append(dist)
  assert np.isfinite(out[0][ccf != 0]).all() and np.array_equal(out[0], out[1])


def test_ball_invalidation(emu):
  """b2t_invalidate_ball (the soma's one-off ball, trace.py:160-168): one seed and several seeds against the oracle's
  literal heap-ordered form -- with one seed every claim order gives the same set."""
  cc, n = _volume(33, shape=(48, 40, 30), n=3)
  an = (16.0, 16.0, 40.0)
  l = int(np.argmax(np.bincount(cc.ravel())[1:]) + 1)
  vol = np.asfortranarray((cc == l).astype(np.uint8))
  dbf = oracle.edt(vol, an, False)
  sx, sy, sz = vol.shape
  ccf = np.ascontiguousarray(vol.reshape(-1, order="F").astype(np.uint32))
  dbff = np.ascontiguousarray(dbf.reshape(-1, order="F"))
  pts = np.argwhere(vol)
  for seeds_xyz, mode in (([tuple(pts[len(pts) // 2])], "heap"), ([tuple(pts[len(pts) // 3]), tuple(pts[2 * len(pts) // 3])], "rounds")):
    ref = vol.copy(order="F")
    n_ref, ref = oracle.roll_invalidation_ball_inside_component(ref, dbf, 6.0, 100.0, an, seeds_xyz, mode=mode)
    claim = np.full(ccf.size, 0xFFFFFFFFFFFFFFFF, np.uint64)
    seeds = np.ascontiguousarray(np.array([x + sx * (y + sy * z) for x, y, z in seeds_xyz], np.uint32))
    nfg = int(vol.sum())
    fv, fs, ctrl = np.zeros(2 * nfg + 8, np.uint32), np.zeros(2 * nfg + 8, np.uint32), np.zeros(16, np.uint32)
    assert emu.b2t_invalidate_ball(p(ccf), p(dbff), p(claim), c_i64(sx), c_i64(sy), c_i64(sz), c_f32(an[0]), c_f32(an[1]),
                                   c_f32(an[2]), p(seeds), c_u32(seeds.size), c_f32(6.0), c_f32(100.0), p(fv), p(fs),
                                   c_u64(nfg), p(ctrl), None) == 0
    got = ((claim != 0) & (ccf == 1)).astype(np.uint8).reshape(vol.shape, order="F")
    assert int(ctrl[6]) == n_ref and np.array_equal(got, ref), (mode, int(ctrl[6]), n_ref)
