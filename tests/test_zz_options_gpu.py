"""Off-default options of skeletonize() on the CUDA path against the oracle.  Written in the
session of round 1 that had no GPU time left: the host logic is covered on CPU tensors in tests/test_oracle_cpu.py,
the kernels involved (b2t_fill_voids, the whole default path) by the files that run before this one."""
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
