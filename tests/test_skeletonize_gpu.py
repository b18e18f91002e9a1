"""End-to-end parity of the CUDA path (kimimaro_b200.skeletonize, through the C ABI) with the CPU
oracle (oracle/teasar.py) on the same inputs: vertex and edge arrays bit-exact, radii within 1e-4
relative (BASELINE.json north_star).  Also the reference's own known-answer tests
(automated_test.py:17-102, 116-199, 261-279) run against the CUDA path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _compare(res, ref, rtol=1e-4):
  assert sorted(res.keys()) == sorted(ref.keys())
  for k in ref:
    a, b = res[k], ref[k]
    assert a.vertices.shape == b["vertices"].shape, (k, a.vertices.shape, b["vertices"].shape)
    assert np.array_equal(a.vertices, b["vertices"]), k
    assert np.array_equal(a.edges, b["edges"]), k
    np.testing.assert_allclose(a.radii, b["radii"], rtol=rtol)


def _both(gpu, labels, **kw):
  import kimimaro_b200
  from oracle import teasar
  res = kimimaro_b200.skeletonize(labels, progress=False, **kw)
  ref = teasar.skeletonize(labels, **kw)
  return res, ref


def test_sphere_config0(gpu):
  from tests.synth import sphere
  res, ref = _both(gpu, sphere(64, 24))
  assert len(res) == 1
  _compare(res, ref)


@pytest.mark.parametrize("seed,shape,n,an", [
  (1, (96, 96, 64), 12, (16, 16, 40)),
  (2, (128, 64, 48), 20, (1, 1, 1)),
  (3, (64, 128, 96), 16, (4, 4, 40)),
])
def test_tubes_small(gpu, seed, shape, n, an):
  from tests.synth import synthetic_tubes
  lab = synthetic_tubes(shape, n, seed=seed, anisotropy=an)
  res, ref = _both(gpu, lab, anisotropy=an, dust_threshold=100)
  assert len(ref) > 0
  _compare(res, ref)


def test_tubes_no_fix_borders(gpu):
  from tests.synth import synthetic_tubes
  lab = synthetic_tubes((96, 96, 64), 10, seed=5)
  res, ref = _both(gpu, lab, anisotropy=(16, 16, 40), dust_threshold=100, fix_borders=False)
  _compare(res, ref)


def test_small_params_many_paths(gpu):
  # small invalidation radius -> many paths per label
  from tests.synth import synthetic_tubes
  lab = synthetic_tubes((96, 96, 64), 6, seed=9)
  tp = {"scale": 1.0, "const": 20, "pdrf_scale": 100000, "pdrf_exponent": 4}
  res, ref = _both(gpu, lab, anisotropy=(16, 16, 40), dust_threshold=100, teasar_params=tp)
  _compare(res, ref)


# ---- the reference's own tests (automated_test.py) against the CUDA path ----
def test_empty_image(gpu):
  import kimimaro_b200 as kimimaro
  assert len(kimimaro.skeletonize(np.zeros((64, 64, 64), dtype=bool), fix_borders=True)) == 0


def test_very_sparse_image(gpu):
  import kimimaro_b200 as kimimaro
  labels = np.zeros((64, 64, 64), dtype=bool)
  labels[5, 5, 5] = True
  labels[6, 5, 5] = True
  labels[20, 20, 20] = True
  assert len(kimimaro.skeletonize(labels, dust_threshold=0)) == 1


def test_solid_image(gpu):
  import kimimaro_b200 as kimimaro
  assert len(kimimaro.skeletonize(np.ones((128, 128, 128), dtype=bool), fix_borders=True)) == 1


def test_square(gpu):
  import kimimaro_b200 as kimimaro
  for corners in (((-1, 0), (0, -1)), ((0, 0), (-1, -1))):
    labels = np.ones((1000, 1000), dtype=np.uint8)
    for c in corners:
      labels[c] = 0
    skels = kimimaro.skeletonize(labels, teasar_params=kimimaro.DEFAULT_TEASAR_PARAMS, fix_borders=False)
    assert len(skels) == 1
    skel = skels[1]
    assert skel.vertices.shape[0] == 1000
    assert skel.edges.shape[0] == 999
    assert abs(skel.cable_length() - 999 * np.sqrt(2)) < 0.001
    assert skel.space == "physical"


def test_cube(gpu):
  import kimimaro_b200 as kimimaro
  labels = np.ones((128, 128, 128), dtype=np.uint8)
  labels[0, 0, 0] = 0
  labels[-1, -1, -1] = 0
  skels = kimimaro.skeletonize(labels, fix_borders=False)
  skel = skels[1]
  assert skel.vertices.shape[0] == 128
  assert skel.edges.shape[0] == 127
  assert abs(skel.cable_length() - 127 * np.sqrt(3)) < 0.001


@pytest.mark.parametrize("axis", ["x", "y", "z"])
def test_fix_borders(gpu, axis):
  import kimimaro_b200 as kimimaro
  labels = np.zeros((256, 256, 256), dtype=np.uint8)
  sl = [slice(64, 196)] * 3
  sl["xyz".index(axis)] = slice(None)
  labels[tuple(sl)] = 128
  an = (40, 32, 20) if axis == "z" else (1, 1, 1)
  skels = kimimaro.skeletonize(labels, teasar_params={"const": 250, "scale": 10, "pdrf_exponent": 4,
                                                     "pdrf_scale": 100000}, anisotropy=an, fix_borders=True)
  skel = skels[128].voxel_space()
  for a in range(3):
    if a == "xyz".index(axis):
      assert np.all(skel.vertices[:, a] == np.arange(256))
    else:
      assert np.all(skel.vertices[:, a] == 129)


def test_dimensions(gpu):
  import kimimaro_b200 as kimimaro
  for shp in ((10,), (10, 10), (10, 10, 10), (10, 10, 10, 1)):
    kimimaro.skeletonize(np.zeros(shp, dtype=np.uint8))
  with pytest.raises(kimimaro.DimensionError):
    kimimaro.skeletonize(np.zeros((10, 10, 10, 2), dtype=np.uint8))


def test_unknown_param_raises(gpu):
  import kimimaro_b200 as kimimaro
  with pytest.raises(TypeError):
    kimimaro.skeletonize(np.ones((32, 32, 32), np.uint8), teasar_params={"bogus": 1}, dust_threshold=0)


def test_golden_fixtures_cuda(gpu):
  """The committed golden vectors (tests/golden/golden_v1.npz) against the CUDA path."""
  import os
  import kimimaro_b200
  from tests.synth import sphere, synthetic_tubes
  g = np.load(__import__("tests.conftest", fromlist=["golden_name"]).golden_name("golden_v1", ".npz"))
  cases = {"sphere": (sphere(64, 24), {}),
           "tubes": (synthetic_tubes((96, 96, 64), 12, seed=1), {"anisotropy": (16, 16, 40), "dust_threshold": 100})}
  for name, (lab, kw) in cases.items():
    sk = kimimaro_b200.skeletonize(lab, progress=False, **kw)
    ids = [int(i) for i in g[f"{name}_ids"]]
    assert sorted(sk) == sorted(ids)
    for i in ids:
      assert np.array_equal(sk[i].vertices, g[f"{name}_{i}_v"])
      assert np.array_equal(sk[i].edges, g[f"{name}_{i}_e"])
      np.testing.assert_allclose(sk[i].radii, g[f"{name}_{i}_r"], rtol=1e-4)


def _blob_volume(hole=True):
  # a ball with dendrites; with the thresholds below it takes the soma branch of trace.py:108-127
  lab = np.zeros((128, 128, 96), np.uint8, order="F")
  xs, ys, zs = np.ogrid[:128, :128, :96]
  lab[(xs - 64) ** 2 + (ys - 64) ** 2 + (zs - 48) ** 2 <= 30 ** 2] = 1
  lab[60:68, 60:68, :] = 1
  lab[:, 62:66, 46:50] = 1
  lab[90:100, 20:120, 40:44] = 1
  if hole:
    lab[60:66, 60:66, 44:50] = 0      # an internal void that fill_voids must close
  return lab


@pytest.mark.parametrize("hole", [False, True])
def test_soma_branch(gpu, hole):
  tp = {"scale": 1.5, "const": 3, "pdrf_scale": 100000, "pdrf_exponent": 4, "soma_detection_threshold": 12,
        "soma_acceptance_threshold": 20, "soma_invalidation_scale": 1.0, "soma_invalidation_const": 2}
  res, ref = _both(gpu, _blob_volume(hole), teasar_params=tp, dust_threshold=100)
  assert len(ref) == 1
  _compare(res, ref)


def test_soma_detected_not_accepted(gpu):
  tp = {"scale": 1.5, "const": 3, "pdrf_scale": 100000, "pdrf_exponent": 4, "soma_detection_threshold": 12,
        "soma_acceptance_threshold": 1000}
  res, ref = _both(gpu, _blob_volume(True), teasar_params=tp, dust_threshold=100)
  _compare(res, ref)


def test_extra_targets_and_object_ids(gpu):
  from tests.synth import synthetic_tubes
  lab = synthetic_tubes((96, 96, 64), 8, seed=21)
  ids = [int(v) for v in np.unique(lab) if v != 0][:4]
  pts = [tuple(int(c) for c in np.argwhere(lab == ids[0])[5])]
  res, ref = _both(gpu, lab, anisotropy=(16, 16, 40), dust_threshold=100, object_ids=ids,
                   extra_targets_after=pts, extra_targets_before=pts)
  _compare(res, ref)


def test_max_paths(gpu):
  from tests.synth import synthetic_tubes
  lab = synthetic_tubes((96, 96, 64), 6, seed=9)
  tp = {"scale": 1.0, "const": 20, "pdrf_scale": 100000, "pdrf_exponent": 4, "max_paths": 3}
  res, ref = _both(gpu, lab, anisotropy=(16, 16, 40), dust_threshold=100, teasar_params=tp)
  _compare(res, ref)


def test_uint64_and_int32_labels(gpu):
  from tests.synth import synthetic_tubes
  lab = synthetic_tubes((64, 64, 48), 5, seed=4)
  big = lab.astype(np.uint64) * np.uint64(2 ** 40 + 7)
  res, ref = _both(gpu, big, anisotropy=(16, 16, 40), dust_threshold=100)
  _compare(res, ref)
  res, ref = _both(gpu, lab.astype(np.int32), anisotropy=(16, 16, 40), dust_threshold=100)
  _compare(res, ref)


def test_ccl_exact_and_deterministic(gpu, orc):
  """N1: the union-find CCL against the oracle's (numbering included), repeated to catch races."""
  import torch
  from kimimaro_b200 import engine, ops
  from tests.synth import synthetic_tubes
  lab = synthetic_tubes((256, 256, 128), 150, seed=33)
  ref, n_ref = orc.connected_components(lab)
  d = ops.to_device_f(lab, gpu).view(torch.int32)
  for _ in range(5):
    cc, n = engine.connected_components(d, lab.shape)
    assert n == n_ref
    got = cc.cpu().numpy().view(np.uint32).reshape(lab.shape, order="F")
    assert np.array_equal(got, ref)


def test_full_size_512_against_oracle_digest(gpu):
  """BASELINE.json's full size (512^3, 1946 labels, soma + glia): every skeleton of the CUDA path must hash
  to the digest of the oracle's result (tests/golden/synth512_oracle_digest.json, produced by
  scripts/full_parity.py), plus size-independent structure checks and run-to-run determinism."""
  import hashlib, json, os
  import kimimaro_b200
  from bench import make_volume, ANISOTROPY
  gold = json.load(open(__import__("tests.conftest", fromlist=["golden_name"]).golden_name("synth512_oracle_digest", ".json")))
  vol = make_volume(512)
  a = kimimaro_b200.skeletonize(vol, anisotropy=ANISOTROPY, progress=False)
  b = kimimaro_b200.skeletonize(vol, anisotropy=ANISOTROPY, progress=False)
  want = gold["sha256_16_of_vertices_then_edges"]
  assert len(a) == gold["n_skeletons"] and sorted(str(k) for k in a) == sorted(want)
  assert sum(s.vertices.shape[0] for s in a.values()) == gold["n_vertices"]
  an = np.array(ANISOTROPY, np.float32)
  bad = []
  for k, s in a.items():
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(s.vertices).tobytes())
    h.update(np.ascontiguousarray(s.edges).tobytes())
    if h.hexdigest()[:16] != want[str(k)]:
      bad.append(k)
    assert np.array_equal(s.vertices, b[k].vertices) and np.array_equal(s.edges, b[k].edges)   # deterministic
    vox = np.rint(s.vertices / an).astype(np.int64)
    assert np.all(vol[vox[:, 0], vox[:, 1], vox[:, 2]] == k)                  # every vertex lies in its label
    d = np.abs(vox[s.edges[:, 0]] - vox[s.edges[:, 1]]).max(axis=1)
    assert np.all(d == 1)                                                     # edges join 26-neighbours
    assert np.all(s.radii > 0)
  assert bad == [], bad[:10]


@pytest.mark.parametrize("seed", [2, 11])
def test_no_fix_branching(gpu, seed):
  """fix_branching=False (trace.py:154-158, 244): one parental field per label, paths by pointer walk."""
  from tests.synth import synthetic_tubes
  lab = synthetic_tubes((96, 96, 64), 8, seed=seed)
  tp = {"scale": 1.0, "const": 30, "pdrf_scale": 100000, "pdrf_exponent": 4}
  res, ref = _both(gpu, lab, anisotropy=(16, 16, 40), dust_threshold=100, teasar_params=tp, fix_branching=False)
  assert len(ref) > 0
  _compare(res, ref)


def test_no_fix_branching_2d_slice(gpu):
  # automated_test.py:562-563 runs fix_branching=False on a 2-D slice
  from tests.synth import synthetic_tubes
  lab = synthetic_tubes((128, 128, 64), 14, seed=3)[:, :, 30]
  res, ref = _both(gpu, lab, anisotropy=(16, 16, 40), dust_threshold=50, fix_branching=False)
  _compare(res, ref)


def test_non_integer_anisotropy(gpu):
  """Anisotropies that are not exactly representable (the x pass multiplies where the library adds repeatedly):
  radii stay within 1e-4, vertices and edges still have to match."""
  from tests.synth import synthetic_tubes
  an = (3.58, 3.58, 4.1)
  lab = synthetic_tubes((96, 96, 64), 10, seed=13, anisotropy=an)
  tp = {"scale": 1.5, "const": 20, "pdrf_scale": 100000, "pdrf_exponent": 4}
  res, ref = _both(gpu, lab, anisotropy=an, dust_threshold=100, teasar_params=tp)
  _compare(res, ref)
