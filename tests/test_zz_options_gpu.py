"""Off-default options of skeletonize() on the CUDA path against the oracle.  Written in the
session of round 1 that had no GPU time left.  Before their first run on a device the bodies of these tests ran on the
CPU, through the whole product, on the library's kernels compiled against the SIMT emulation
(`python scripts/run_gpu_tests_emulated.py tests.test_zz_options_gpu <name>`): test_fill_holes, test_fix_avocados,
test_fill_all_holes_reference_known_answer, test_binary_image and test_extra_targets_grow_the_skeleton passed there; the
two with million-voxel labels (test_square_with_fill_holes, test_parallel_argument_is_accepted) are too slow to emulate
and only add fill_holes=True / parallel=2 to cases the GPU suite already runs.  Scaled-down versions are part of the CPU
suite (tests/test_product_on_emulated_library_cpu.py)."""
import numpy as np
import pytest

from tests.test_skeletonize_gpu import _both, _compare

pytestmark = pytest.mark.gpu


def _cell_with_nucleus():
  v = np.zeros((96, 80, 64), np.uint32, order="F")
  v[8:88, 20:60, 12:52] = 5          # a cell body ...
  v[30:50, 30:50, 24:40] = 9         # ... whose nucleus is a label of its own
  v[60:70, 34:44, 28:36] = 0         # ... and a vacuole (an enclosed void)
  v[8:88, 66:76, 20:40] = 7          # a second, solid process
  v[40:44, 60:66, 28:32] = 7         # touching nothing: separate component of 7? no -- attached to it
  return v


def test_fill_holes(gpu):
  labels = _cell_with_nucleus()
  kw = dict(anisotropy=(16, 16, 40), dust_threshold=100, fill_holes=True)
  res, ref = _both(gpu, labels, **kw)
  assert 9 not in ref and 5 in ref and 7 in ref      # the nucleus was swallowed (intake.py:787-790)
  _compare(res, ref)
  # without the option the nucleus is an object of its own and the cell is hollow: a different answer
  res0, ref0 = _both(gpu, labels, anisotropy=(16, 16, 40), dust_threshold=100)
  assert 9 in ref0
  _compare(res0, ref0)
  assert not np.array_equal(res0[5].vertices, res[5].vertices)


def test_fix_avocados(gpu):
  import kimimaro_b200
  labels = _cell_with_nucleus()
  params = dict(kimimaro_b200.DEFAULT_TEASAR_PARAMS)
  params["soma_detection_threshold"] = 300        # candidates: DBF above 300 / 2.5 nm (intake.py:619)
  kw = dict(anisotropy=(16, 16, 40), dust_threshold=100, fix_avocados=True, teasar_params=params)
  res, ref = _both(gpu, labels, **kw)
  assert 9 not in ref and 5 in ref and 7 in ref    # the nucleus took the label of the cell around it
  _compare(res, ref)
  both = dict(kw, fill_holes=True)
  res2, ref2 = _both(gpu, labels, **both)
  _compare(res2, ref2)


# ---- the reference's own known-answer tests that the options above unlock, and a few it has besides ----
def test_square_with_fill_holes(gpu):
  """automated_test.py:49-87 with fill_holes=True (the reference parametrises test_square over it)."""
  import kimimaro_b200 as kimimaro
  for corners in (((-1, 0), (0, -1)), ((0, 0), (-1, -1))):
    labels = np.ones((1000, 1000), dtype=np.uint8)
    for c in corners:
      labels[c] = 0
    skels = kimimaro.skeletonize(labels, teasar_params=kimimaro.DEFAULT_TEASAR_PARAMS, fix_borders=False, fill_holes=True,
                                 progress=False)
    assert len(skels) == 1
    skel = skels[1]
    assert skel.vertices.shape[0] == 1000 and skel.edges.shape[0] == 999
    assert abs(skel.cable_length() - 999 * np.sqrt(2)) < 0.001


def test_fill_all_holes_reference_known_answer(gpu):
  """automated_test.py:458-476: two solid halves full of foreign-label noise; afterwards only the halves remain."""
  import torch
  from kimimaro_b200 import engine, intake, ops
  rng = np.random.default_rng(1)
  labels = np.zeros((64, 32, 32), dtype=np.uint32, order="F")
  labels[0:32] = 1
  labels[32:64] = 8
  labels[1:31, 1:31, 1:31] = rng.integers(1, 8, size=(30, 30, 30))
  labels[33:63, 1:31, 1:31] = rng.integers(8, 11, size=(30, 30, 30))
  assert set(np.unique(labels)) == set(range(1, 11))
  d = ops.to_device_f(labels, gpu).view(torch.int32)
  V = labels.size
  count, bbox, _, _ = engine.label_stats(d, torch.zeros(V, dtype=torch.float32, device=d.device), labels.shape, 10)
  intake.fill_all_holes(d, labels.shape, 10, count.cpu().numpy(), bbox.cpu().numpy().reshape(-1, 6))
  assert set(torch.unique(d).cpu().tolist()) == {1, 8}


def test_binary_image(gpu):
  """automated_test.py:39-46."""
  import kimimaro_b200 as kimimaro
  labels = np.ones((256, 256, 3), dtype=bool)
  labels[-1, 0] = 0
  labels[0, -1] = 0
  assert len(kimimaro.skeletonize(labels, fix_borders=False, progress=False)) == 1


def test_extra_targets_grow_the_skeleton(gpu):
  """automated_test.py:201-232."""
  import kimimaro_b200 as kimimaro
  labels = np.zeros((256, 256, 1), dtype=np.uint8)
  labels[64:196, 64:196, :] = 128
  tp = {"const": 250, "scale": 10, "pdrf_exponent": 4, "pdrf_scale": 100000}

  def run(**kw):
    return kimimaro.skeletonize(labels, teasar_params=tp, anisotropy=(1, 1, 1), dust_threshold=1000, progress=False,
                                fix_branching=True, fix_borders=True, **kw)[128]
  skel1 = run()
  skel2 = run(extra_targets_after=[(65, 65, 0)])
  skel3 = run(extra_targets_before=[(65, 65, 0)])
  assert skel1.vertices.size < skel2.vertices.size
  assert skel3.vertices.size < skel2.vertices.size


def test_parallel_argument_is_accepted(gpu):
  """automated_test.py:234-259: four quadrants, parallel=2 (here: ignored, one GPU traces all labels)."""
  import kimimaro_b200 as kimimaro
  labels = np.zeros((256, 256, 128), dtype=np.uint8)
  labels[0:128, 0:128, :] = 1
  labels[0:128, 128:256, :] = 2
  labels[128:256, 0:128, :] = 3
  labels[128:256, 128:256, :] = 4
  tp = {"const": 250, "scale": 10, "pdrf_exponent": 4, "pdrf_scale": 100000}
  skels = kimimaro.skeletonize(labels, teasar_params=tp, anisotropy=(1, 1, 1), dust_threshold=1000, progress=False,
                               parallel=2)
  assert len(skels) == 4


def test_connect_points(gpu):
  """kimimaro.connect_points (intake.py:268-313, trace.py:358-390) on the CUDA path against the oracle's restatement:
  vertices in path order (end first), radii, both anisotropies; disconnected end points raise."""
  import kimimaro_b200
  from oracle import teasar
  from tests.synth import synthetic_tubes
  lab = synthetic_tubes((96, 80, 64), 4, seed=9)
  ids = [int(v) for v in np.unique(lab) if v]
  for an in ((16, 16, 40), (1, 1, 1), (4, 4, 40)):
    for seg in ids[:2]:
      pts = np.argwhere(lab == seg)
      start, end = tuple(int(v) for v in pts[0]), tuple(int(v) for v in pts[-1])
      got = kimimaro_b200.connect_points(lab == seg, start, end, anisotropy=an)
      ref = teasar.connect_points(lab == seg, start, end, anisotropy=an)
      assert np.array_equal(got.vertices, ref["vertices"]) and np.array_equal(got.edges, ref["edges"])
      np.testing.assert_allclose(got.radii, ref["radii"], rtol=1e-4)
  with pytest.raises(ValueError):                             # a background voxel belongs to no component
    kimimaro_b200.connect_points(lab, tuple(int(v) for v in np.argwhere(lab == 0)[0]),
                                 tuple(int(v) for v in np.argwhere(lab == ids[0])[0]))


@pytest.mark.parametrize("axis", ["x", "y"])
def test_joinability_and_postprocess(gpu, axis):
  """automated_test.py:282-333: two chunks that overlap in one z plane; with fix_borders=True their skeletons meet in that
  plane and merge into one component, without it they do not give the same merged skeleton.  Then the chunk-stitch
  post-processing (kimimaro_b200.postprocess, post.py:49-87) leaves one cycle-free component with the label's id."""
  import kimimaro_b200 as kimimaro
  from kimimaro_b200 import post
  from kimimaro_b200.skeleton import Skeleton

  def run(labels, fix_borders):
    return kimimaro.skeletonize(
      labels, teasar_params={"const": 10, "scale": 10, "pdrf_exponent": 4, "pdrf_scale": 100000}, anisotropy=(1, 1, 1),
      object_ids=None, dust_threshold=0, progress=False, fix_branching=True, in_place=False, fix_borders=fix_borders,
      parallel=1)

  labels = np.zeros((256, 256, 20), dtype=np.uint8)
  labels[(np.s_[32:160, :, :] if axis == "x" else np.s_[:, 32:160, :])] = 1

  def halves(fix_borders):
    a = run(labels[:, :, :10], fix_borders)[1]
    b = run(labels[:, :, 9:], fix_borders)[1]
    b.vertices[:, 2] += 9
    return a.merge(b)

  merged_fb = halves(True)
  assert len(merged_fb.components()) == 1
  merged = halves(False)
  assert not Skeleton.equivalent(merged, merged_fb)

  out = kimimaro.postprocess(merged_fb, dust_threshold=0, tick_threshold=5)
  assert out.id == merged_fb.id
  assert len(out.components()) == 1
  assert len(post.find_cycle(out.edges.astype(np.int32))) == 0
  assert out.edges.shape[0] == out.vertices.shape[0] - 1        # a tree
