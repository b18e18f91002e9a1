"""The PRODUCT on the CPU suite: kimimaro_b200.skeletonize -- its whole Python host path (intake.py, engine.py,
border.py, skeleton.py) -- runs here on CPU tensors against oracle/_cache/libb2t_emu.so, the library's own kernels
compiled for the CPU against the SIMT emulation (tests/host/*_emu.cpp; K1 through the host context of its column-pass
templates), and must give the oracle's skeletons bit for bit, like the GPU tests ask of the real library.

This is test infrastructure: the product is not touched -- the test swaps the library path and stubs the handful of
torch.cuda calls the host code makes (tensor.cuda(), current_stream, is_available); without those stubs the product
still refuses to run without an sm_100 device (tests/test_oracle_cpu.py::test_no_cpu_fallback)."""
import contextlib
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HOST = os.path.join(HERE, "host")
UNITS = ["trace_emu", "preamble_emu", "field_emu", "b2t_emu_main"]
OUT = os.path.join(ROOT, "oracle", "_cache", "libb2t_emu.so")


def _build():
  deps = [os.path.join(HOST, u + ".cpp") for u in UNITS] + [os.path.join(HOST, "fh3_host.cpp")]
  deps += [os.path.join(HOST, "emu_include", f) for f in os.listdir(os.path.join(HOST, "emu_include"))]
  deps += [os.path.join(ROOT, "kimimaro_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "kimimaro_b200", "csrc"))]
  if os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in deps):
    return OUT
  obj_dir = os.path.join(ROOT, "oracle", "_cache", "emuobj")
  os.makedirs(obj_dir, exist_ok=True)
  flags = ["-O1", "-ffp-contract=off", "-std=c++17", "-fPIC", "-pthread", "-Wno-attributes", "-DB2T_EMU_COMBINED",
           "-I" + os.path.join(HOST, "emu_include"), "-I" + HOST]
  objs = []
  for u in UNITS:
    o = os.path.join(obj_dir, u + ".o")
    subprocess.check_call(["g++"] + flags + ["-c", os.path.join(HOST, u + ".cpp"), "-o", o])
    objs.append(o)
  subprocess.check_call(["g++", "-shared", "-pthread"] + objs + ["-o", OUT])
  return OUT


class _Stream:
  cuda_stream = 0


def _emulate(setattr_, lib_path):
  """Point kimimaro_b200 at the emulated library and stub the torch.cuda calls its host code makes."""
  import torch
  from kimimaro_b200 import _lib
  import kimimaro_b200.intake  # noqa: F401  (registers its entry points with _lib.declare)
  import kimimaro_b200.engine  # noqa: F401
  setattr_(_lib, "LIB_PATH", lib_path)
  setattr_(_lib, "_lib", None)
  setattr_(_lib, "_checked_devices", set())
  setattr_(torch.cuda, "is_available", lambda: True)
  setattr_(torch.cuda, "current_device", lambda: 0)
  setattr_(torch.cuda, "current_stream", lambda *a, **k: _Stream())
  setattr_(torch.cuda, "synchronize", lambda *a, **k: None)
  setattr_(torch.cuda, "stream", lambda s: contextlib.nullcontext())
  setattr_(torch.Tensor, "cuda", lambda self, *a, **k: self)
  setattr_(torch.Tensor, "is_cuda", property(lambda self: True))   # ops.edt asserts it
  import kimimaro_b200
  return kimimaro_b200


@pytest.fixture()
def product(monkeypatch):
  yield _emulate(lambda o, n, v: monkeypatch.setattr(o, n, v, raising=False), _build())
  monkeypatch.undo()


def _compare(res, ref):
  assert sorted(res.keys()) == sorted(ref.keys())
  for k in ref:
    a, b = res[k], ref[k]
    assert np.array_equal(a.vertices, b["vertices"]), k
    assert np.array_equal(a.edges, b["edges"]), k
    np.testing.assert_allclose(a.radii, b["radii"], rtol=1e-4)


def test_sphere_and_tubes(product):
  from oracle import teasar
  from tests.synth import sphere, synthetic_tubes
  for labels, kw in ((sphere(40, 14), {"dust_threshold": 100}),
                     (synthetic_tubes((56, 48, 32), 5, seed=3), {"anisotropy": (16, 16, 40), "dust_threshold": 100})):
    res = product.skeletonize(labels, progress=False, **kw)
    ref = teasar.skeletonize(labels, **kw)
    assert len(ref) >= 1
    _compare(res, ref)


def _small_cell_with_nucleus():
  v = np.zeros((48, 40, 32), np.uint32, order="F")
  v[4:44, 10:30, 6:26] = 5           # a cell body ...
  v[15:25, 15:25, 12:20] = 9         # ... whose nucleus is a label of its own
  v[30:35, 17:22, 14:18] = 0         # ... and a vacuole (an enclosed void)
  v[4:44, 33:38, 10:20] = 7          # a second, solid process
  return v


def test_fill_holes_and_fix_avocados(product):
  """The options of SURVEY 8f N4 through the product's real pipeline (tests/test_zz_options_gpu.py asks the same of the
  GPU): nucleus swallowed, vacuole filled, skeletons equal to the oracle's."""
  from oracle import teasar
  labels = _small_cell_with_nucleus()
  base = dict(anisotropy=(16, 16, 40), dust_threshold=50)
  ref0 = teasar.skeletonize(labels, **base)
  assert sorted(ref0) == [5, 7, 9]
  _compare(product.skeletonize(labels, progress=False, **base), ref0)
  kw = dict(base, fill_holes=True)
  ref = teasar.skeletonize(labels, **kw)
  assert sorted(ref) == [5, 7]
  _compare(product.skeletonize(labels, progress=False, **kw), ref)
  params = dict(product.DEFAULT_TEASAR_PARAMS)
  params["soma_detection_threshold"] = 150        # candidates: DBF above 150 / 2.5 nm (intake.py:619)
  kw = dict(base, fix_avocados=True, teasar_params=params)
  ref = teasar.skeletonize(labels, **kw)
  assert sorted(ref) == [5, 7]
  _compare(product.skeletonize(labels, progress=False, **kw), ref)
  kw["fill_holes"] = True
  _compare(product.skeletonize(labels, progress=False, **kw), teasar.skeletonize(labels, **kw))


def test_claim_orders_end_to_end(product):
  """The three claim orders of the path loop (b2t_trace_batch: invalidation_mode) through the whole product: the default
  (key-ordered rounds of one voxel) against oracle mode 'window:1', the strict mode against oracle mode 'heap' -- which
  is the reference's compiled extension voxel for voxel -- and the hop-synchronous rounds against 'rounds'."""
  from kimimaro_b200 import _lib
  from oracle import teasar
  from tests.synth import synthetic_tubes
  labels = synthetic_tubes((56, 48, 40), 6, seed=9)
  kw = dict(anisotropy=(16, 16, 40), dust_threshold=100, teasar_params=dict(product.DEFAULT_TEASAR_PARAMS, scale=1.5, const=30))
  assert _lib.invalidation_mode() == ("window", 1.0)
  _compare(product.skeletonize(labels, progress=False, **kw), teasar.skeletonize(labels, invalidation_mode="window:1", **kw))
  try:
    for mode, omode in (("strict", "heap"), ("rounds", "rounds")):
      _lib.set_invalidation_mode(mode)
      _compare(product.skeletonize(labels, progress=False, **kw), teasar.skeletonize(labels, invalidation_mode=omode, **kw))
  finally:
    _lib.set_invalidation_mode("window", 1.0)


def test_path_pool_bounds(product):
  """(a) fix_branching=True with small radii on a slab: every voxel is its own path, 3 slots each -- the
  pool bound is 3 * n_fg + 2 * targets; (b) fix_branching=False on a comb: every tooth's path is the whole root -> tip
  walk, far beyond any multiple of n_fg that is affordable up front -- the arena is traced again with a larger pool."""
  from oracle import teasar
  slab = np.zeros((38, 38, 6), np.uint8, order="F")
  slab[1:37, 1:37, 1:5] = 1
  for tp in (dict(scale=0.5, const=0), dict(scale=0, const=1)):
    kw = dict(dust_threshold=10, teasar_params=dict(product.DEFAULT_TEASAR_PARAMS, **tp))
    _compare(product.skeletonize(slab, progress=False, **kw), teasar.skeletonize(slab, **kw))
  comb = np.zeros((120, 44, 3), np.uint8, order="F")
  comb[1:119, 1:3, 1] = 1                                             # the spine
  for x in range(2, 118, 2):
    comb[x, 3:42, 1] = 1                                              # 58 one-voxel-thin teeth
  kw = dict(dust_threshold=10, fix_branching=False, fix_borders=False,
            teasar_params=dict(product.DEFAULT_TEASAR_PARAMS, scale=1.5, const=2))
  _compare(product.skeletonize(comb, progress=False, **kw), teasar.skeletonize(comb, **kw))


def test_more_of_the_api(product):
  """Soma branch (detected and accepted: fill, re-EDT, free-space DAF, ball invalidation, cull), extra targets,
  object_ids, fix_branching=False, max_paths, a 2-D image, uint64 labels."""
  from oracle import teasar
  from tests.synth import synthetic_tubes
  ball = np.zeros((44, 44, 30), np.uint32, order="F")
  x, y, z = np.ogrid[:44, :44, :30]
  ball[(x - 22) ** 2 + (y - 22) ** 2 + ((z - 15) * 2.5) ** 2 <= 15 ** 2] = 3
  ball[20:24, 20:24, :] = 3                                           # a process through the soma
  ball[21:23, 21:23, 14:16] = 0                                       # and a void in its middle
  tp = dict(product.DEFAULT_TEASAR_PARAMS, soma_detection_threshold=100, soma_acceptance_threshold=180)
  kw = dict(anisotropy=(16, 16, 40), dust_threshold=100, teasar_params=tp)
  _compare(product.skeletonize(ball, progress=False, **kw), teasar.skeletonize(ball, **kw))
  # ... the default soma ball (2 * DBF + 300 nm) swallows this whole label; with a small one paths remain (root at the soma
  # centre, cull of path points inside the soma radius), and with acceptance out of reach the private arena traces it as an
  # ordinary label
  for extra in (dict(soma_invalidation_scale=0.5, soma_invalidation_const=0), dict(soma_acceptance_threshold=1e9)):
    kw2 = dict(kw, teasar_params=dict(tp, **extra))
    ref = teasar.skeletonize(ball, **kw2)
    assert 3 in ref and ref[3]["vertices"].shape[0] > 10
    _compare(product.skeletonize(ball, progress=False, **kw2), ref)
  tubes = synthetic_tubes((56, 48, 32), 5, seed=4)
  ids = [int(v) for v in np.unique(tubes) if v][:3]
  pt = tuple(int(v) for v in np.argwhere(tubes == ids[0])[7])
  for extra in (dict(object_ids=ids[:2]), dict(extra_targets_after=[pt]), dict(extra_targets_before=[pt]),
                dict(fix_branching=False), dict(fix_borders=False), dict(teasar_params=dict(product.DEFAULT_TEASAR_PARAMS, max_paths=2))):
    kw = dict(dict(anisotropy=(16, 16, 40), dust_threshold=100), **extra)
    _compare(product.skeletonize(tubes, progress=False, **kw), teasar.skeletonize(tubes, **kw))
  kw = dict(anisotropy=(16, 16, 40), dust_threshold=100)
  _compare(product.skeletonize(tubes.astype(np.uint64) * 10 ** 10, progress=False, **kw),
           {k * 10 ** 10: v for k, v in teasar.skeletonize(tubes, **kw).items()})
  plane = np.asfortranarray(tubes[:, :, 16])
  kw = dict(anisotropy=(16, 16), dust_threshold=20) if False else dict(dust_threshold=20)
  _compare(product.skeletonize(plane, progress=False, **kw), teasar.skeletonize(plane, **kw))


@pytest.mark.parametrize("axis", ["x", "y"])
def test_joinability_and_postprocess(product, axis):
  """automated_test.py:282-333 scaled down (64 x 64 x 20 instead of 256 x 256 x 20), then the chunk-stitch post-processing:
  chunks overlapping in one z plane merge into ONE component when fix_borders=True put their end points on the same
  face voxels; kimimaro_b200.postprocess leaves a single cycle-free tree.  Every chunk result equals the oracle's."""
  from kimimaro_b200 import post
  from kimimaro_b200.skeleton import Skeleton
  from oracle import teasar
  tp = {"const": 10, "scale": 10, "pdrf_exponent": 4, "pdrf_scale": 100000}
  labels = np.zeros((64, 64, 20), dtype=np.uint8)
  labels[(np.s_[8:40, :, :] if axis == "x" else np.s_[:, 8:40, :])] = 1

  def halves(fix_borders):
    kw = dict(teasar_params=tp, anisotropy=(1, 1, 1), dust_threshold=0, fix_borders=fix_borders)
    parts = []
    for chunk in (labels[:, :, :10], labels[:, :, 9:]):
      res = product.skeletonize(chunk, progress=False, parallel=1, **kw)
      _compare(res, teasar.skeletonize(chunk, **kw))
      parts.append(res[1])
    parts[1].vertices[:, 2] += 9
    return parts[0].merge(parts[1])

  merged_fb = halves(True)
  assert len(merged_fb.components()) == 1
  assert not Skeleton.equivalent(halves(False), merged_fb)
  out = product.postprocess(merged_fb, dust_threshold=0, tick_threshold=5)
  assert out.id == merged_fb.id
  assert len(out.components()) == 1
  assert len(post.find_cycle(out.edges.astype(np.int32))) == 0
  assert out.edges.shape[0] == out.vertices.shape[0] - 1


# ---- N > 1: labels sharded over two ranks (gloo), one gather of packed skeleton buffers to rank 0 ----
def _soma_and_tubes():
  """a soma (detected and accepted with the thresholds below: private arena) next to two ordinary tubes"""
  from tests.synth import synthetic_tubes
  vol = np.zeros((76, 44, 30), np.uint32, order="F")
  x, y, z = np.ogrid[:44, :44, :30]
  ball = np.zeros((44, 44, 30), np.uint32)
  ball[(x - 22) ** 2 + (y - 22) ** 2 + ((z - 15) * 2.5) ** 2 <= 15 ** 2] = 3
  ball[20:24, 20:24, :] = 3
  ball[21:23, 21:23, 14:16] = 0
  vol[:44] = ball
  tubes = synthetic_tubes((30, 44, 30), 2, seed=8)
  vol[46:] = np.where(tubes != 0, tubes + 10, 0)
  # accepted as a soma, with a small one-off ball so that paths remain (the default 2 * DBF + 300 nm swallows the whole label)
  tp = dict(soma_detection_threshold=100, soma_acceptance_threshold=180, soma_invalidation_scale=0.5, soma_invalidation_const=0)
  return vol, tp


def _sharded_worker(rank, world, port, lib_path, q, with_soma=False):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  import torch
  import torch.distributed as dist
  product = _emulate(setattr, lib_path)
  from kimimaro_b200 import distributed as kd
  from tests.synth import synthetic_tubes
  dist.init_process_group("gloo", rank=rank, world_size=world)
  labels = synthetic_tubes((56, 48, 32), 6, seed=12)
  extra = {}
  if with_soma:
    labels, tp = _soma_and_tubes()
    extra["teasar_params"] = dict(product.DEFAULT_TEASAR_PARAMS, **tp)
  # skeletonize_sharded (kimimaro_b200/distributed.py) as it runs on GPUs, on CPU tensors: every rank uploads its piece of
  # the volume, all-gather, LPT share of the components, ONE gather of the raw path buffers, one assembly on rank 0
  tm = {}
  out = kd.skeletonize_sharded(labels, device=torch.device("cpu"), anisotropy=(16, 16, 40), dust_threshold=100,
                               progress=False, timings=tm, **extra)
  if rank == 0:
    q.put(({k: (v.vertices.copy(), v.edges.copy(), v.radii.copy()) for k, v in out.items()}, tm["n_traced"]))
  else:
    assert out is None
  dist.barrier()
  dist.destroy_process_group()


def test_sharded_over_two_ranks_gloo():
  """BASELINE.json configs[3] in miniature: every rank runs the replicated preamble and traces its LPT share of the
  connected components; rank 0 ends up with exactly the skeletons one process computes."""
  import torch.multiprocessing as mp
  from oracle import teasar
  from tests.synth import synthetic_tubes
  from tests.test_distributed_cpu import _free_port
  lib_path = _build()
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, lib_path, q)) for r in range(2)]
  for p in procs:
    p.start()
  out, n_rank0 = q.get(timeout=600)
  for p in procs:
    p.join(timeout=600)
    assert p.exitcode == 0
  ref = teasar.skeletonize(synthetic_tubes((56, 48, 32), 6, seed=12), anisotropy=(16, 16, 40), dust_threshold=100)
  assert sorted(out) == sorted(ref) and 0 < n_rank0 < len(ref)      # rank 0 traced only a share
  for k in ref:
    assert np.array_equal(out[k][0], ref[k]["vertices"]) and np.array_equal(out[k][1], ref[k]["edges"])
    np.testing.assert_allclose(out[k][2], ref[k]["radii"], rtol=1e-4)


def test_assemble_kernel_equals_general_path(product):
  """b2t_assemble (one CTA per label: dense-volume hash set + two shared-memory bitonic sorts) against the torch
  sort / unique formulation of the same consolidate semantics, on random path buffers: several segments per group,
  revisited voxels, repeated edges, a path of one voxel (dropped: no edge), segments of one group far apart in the
  buffer, a private-arena group, and one group above the kernel's capacity (general path for that group only)."""
  import torch
  from kimimaro_b200 import engine
  rng = np.random.default_rng(11)
  shape = (40, 36, 30)
  V = shape[0] * shape[1] * shape[2]
  cap = int(engine.lib().b2t_assemble_group_cap())
  vox, gids, priv, lens = [], [], [], []

  def walk(n):
    p = rng.integers(5, 25, size=3)
    out = []
    for _ in range(n):
      p = np.clip(p + rng.integers(-1, 2, size=3), 0, np.array(shape) - 1)
      out.append(int(p[0] + shape[0] * (p[1] + shape[1] * p[2])))
    return out

  for seg in range(60):
    g = int(rng.integers(1, 12))
    entries = []
    for _ in range(int(rng.integers(1, 5))):
      entries += walk(int(rng.integers(1, 40))) + [-1]
    vox += entries
    lens.append(len(entries))
    gids.append(g)
    priv.append(g == 7)
  big = []                                                      # group 99: more entries than one CTA takes
  while len(big) <= cap:
    big += walk(200) + [-1]
  vox += big
  lens.append(len(big))
  gids.append(99)
  priv.append(False)
  d_vox = torch.tensor(vox, dtype=torch.int32)
  d_rad = torch.tensor(rng.random(len(vox)).astype(np.float32))
  seg_off = np.concatenate(([0], np.cumsum(lens)))
  seg_ids = np.arange(len(lens))
  kw = dict(shape=shape, anisotropy=(16.0, 16.0, 40.0), offset=(3, 0, 1), group_ids=np.array(gids))
  got = engine.assemble(d_vox, d_rad, seg_off, seg_ids, seg_private=np.array(priv), **kw)
  ref = engine._assemble_general(d_vox, d_rad, seg_off, seg_ids, **kw)
  assert sorted(got) == sorted(ref) and 99 in got
  for k in ref:
    assert np.array_equal(got[k][0], ref[k][0]) and np.array_equal(got[k][1], ref[k][1]), k
    # radii: the general path takes the first occurrence too
    assert np.array_equal(got[k][2], ref[k][2]), k


def test_assemble_kernel_edge_cases(product):
  """b2t_assemble at its edges: a group of exactly the kernel's capacity, a group whose only path has one voxel (no edge:
  the label yields nothing), a path that returns to a voxel (cycle edge kept once), two segments of one group sharing
  their junction voxel, and an empty buffer."""
  import torch
  from kimimaro_b200 import engine
  shape = (64, 64, 64)
  cap = int(engine.lib().b2t_assemble_group_cap())
  kw = dict(shape=shape, anisotropy=(4.0, 4.0, 40.0))
  assert engine.assemble(torch.zeros(0, dtype=torch.int32), torch.zeros(0), np.zeros(1, np.int64), np.zeros(0, np.int64), **kw) == {}
  lin = lambda x, y, z: x + 64 * (y + 64 * z)
  segs = {
    1: [lin(5, 5, 5), -1],                                                  # single voxel: dropped
    2: [lin(1, 1, 1), lin(2, 1, 1), lin(2, 2, 1), lin(1, 1, 1), -1],        # returns to its start
    3: [lin(9, 9, 9), lin(10, 9, 9), -1],
    4: [lin(10, 9, 9), lin(10, 10, 9), -1],                                 # same group as 3 below: shared junction
  }
  full = []                                                                 # exactly `cap` entries: a snake through the volume
  x = y = z = 20
  while len(full) < cap - 1:
    full.append(lin(x, y, z))
    x += 1
    if x == 60:
      x, y = 20, y + 1
      if y == 60:
        y, z = 20, z + 1
  full.append(-1)
  assert len(full) == cap
  segs[5] = full
  gid = {1: 11, 2: 12, 3: 13, 4: 13, 5: 15}
  vox, lens, gids = [], [], []
  for k in sorted(segs):
    vox += segs[k]
    lens.append(len(segs[k]))
    gids.append(gid[k])
  d_vox = torch.tensor(vox, dtype=torch.int32)
  d_rad = torch.arange(len(vox), dtype=torch.float32)
  seg_off = np.concatenate(([0], np.cumsum(lens)))
  got = engine.assemble(d_vox, d_rad, seg_off, np.arange(len(lens)), group_ids=np.array(gids), **kw)
  ref = engine._assemble_general(d_vox, d_rad, seg_off, np.arange(len(lens)), group_ids=np.array(gids), **kw)
  assert sorted(got) == sorted(ref) == [12, 13, 15]
  for k in ref:
    assert all(np.array_equal(a, b) for a, b in zip(got[k], ref[k])), k
  assert got[12][0].shape == (3, 3) and got[12][1].shape == (3, 2)          # triangle: three vertices, three edges
  assert got[13][0].shape == (3, 3) and got[13][1].shape == (2, 2)
  assert got[15][0].shape == (cap - 1, 3)


def test_c_order_input_equals_fortran_input(product):
  """A C-contiguous volume is uploaded as it lies and transposed on the device; a Fortran-contiguous one is uploaded
  without the reference's defensive host copy (the host array is never written).  Same skeletons either way, and the
  caller's arrays are untouched."""
  from tests.synth import synthetic_tubes
  lab_f = np.asfortranarray(synthetic_tubes((44, 36, 28), 4, seed=21))
  lab_c = np.ascontiguousarray(lab_f)
  assert lab_c.flags["C_CONTIGUOUS"] and not lab_c.flags["F_CONTIGUOUS"]
  keep_f, keep_c = lab_f.copy(), lab_c.copy()
  kw = dict(anisotropy=(16, 16, 40), dust_threshold=50, progress=False)
  a = product.skeletonize(lab_f, **kw)
  b = product.skeletonize(lab_c, **kw)
  assert sorted(a) == sorted(b) and len(a) > 0
  for k in a:
    assert np.array_equal(a[k].vertices, b[k].vertices) and np.array_equal(a[k].edges, b[k].edges)
    assert np.array_equal(a[k].radii, b[k].radii)
  assert np.array_equal(lab_f, keep_f) and np.array_equal(lab_c, keep_c)


def test_connect_points(product):
  """kimimaro.connect_points (intake.py:268-313 -> trace.point_to_point, trace.py:358-390): the product (DAF from start,
  PDRF, node-weighted field from end, parent walk: one path through the fix_branching=False machinery) against the
  oracle's restatement, vertices in path order; two voxels in different components raise like the reference."""
  from oracle import teasar
  from tests.synth import synthetic_tubes
  lab = synthetic_tubes((48, 40, 32), 3, seed=5)
  for an in ((16, 16, 40), (1, 1, 1)):
    ids = [int(v) for v in np.unique(lab) if v]
    pts = np.argwhere(lab == ids[0])
    start, end = tuple(int(v) for v in pts[0]), tuple(int(v) for v in pts[-1])
    got = product.connect_points(lab, start, end, anisotropy=an)
    ref = teasar.connect_points(lab, start, end, anisotropy=an)
    assert np.array_equal(got.vertices, ref["vertices"]) and np.array_equal(got.edges, ref["edges"])
    np.testing.assert_allclose(got.radii, ref["radii"], rtol=1e-4)
    assert got.space == "physical" and got.vertices.shape[0] >= 2
    assert np.array_equal(got.vertices[0], np.array(end, np.float32) * np.array(an, np.float32))    # the path starts at `end`
    assert np.array_equal(got.vertices[-1], np.array(start, np.float32) * np.array(an, np.float32))
  with pytest.raises(ValueError):
    product.connect_points(lab, tuple(int(v) for v in np.argwhere(lab == 0)[0]), end)


def test_sharded_with_a_private_arena_gloo():
  """Two gloo ranks, one of which traces the soma label in its private arena: the raw path buffers of both (the private
  flag travels with the segments) are assembled on rank 0 in one pass and equal the oracle's skeletons."""
  import torch.multiprocessing as mp
  from oracle import teasar
  from tests.test_distributed_cpu import _free_port
  lib_path = _build()
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, lib_path, q, True)) for r in range(2)]
  for p in procs:
    p.start()
  out, _ = q.get(timeout=600)
  for p in procs:
    p.join(timeout=600)
    assert p.exitcode == 0
  vol, tp = _soma_and_tubes()
  ref = teasar.skeletonize(vol, anisotropy=(16, 16, 40), dust_threshold=100,
                           teasar_params=dict(teasar.DEFAULT_TEASAR_PARAMS, **tp))
  assert sorted(out) == sorted(ref) and 3 in ref and len(ref) >= 2
  for k in ref:
    assert np.array_equal(out[k][0], ref[k]["vertices"]) and np.array_equal(out[k][1], ref[k]["edges"]), k
    np.testing.assert_allclose(out[k][2], ref[k]["radii"], rtol=1e-4)
