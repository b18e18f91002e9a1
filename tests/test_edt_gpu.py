"""K1 parity: b2t_edt (CUDA, through the C ABI) vs the CPU oracle on the same seeded inputs.
Tolerance: 1e-4 relative (BASELINE.json north_star: 'within 1e-4 relative on EDT/radius floats')."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _run(gpu, orc, lab, an, bb):
  from kimimaro_b200 import ops
  d = ops.to_device_f(lab, gpu)
  out = ops.to_host_f(ops.edt(d, lab.shape, an, bb), lab.shape)
  ref = orc.edt(lab, an, bb)
  assert out.shape == ref.shape
  inf_o, inf_r = np.isinf(out), np.isinf(ref)
  assert np.array_equal(inf_o, inf_r)
  assert np.array_equal(out == 0, ref == 0)
  fin = ~inf_r
  np.testing.assert_allclose(out[fin], ref[fin], rtol=RTOL, atol=0)
  return out, ref


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.uint32, np.uint64])
@pytest.mark.parametrize("bb", [False, True])
def test_random_multilabel(gpu, orc, dtype, bb):
  rng = np.random.default_rng(11)
  for shape, an in [((37, 29, 23), (1, 1, 1)), ((64, 64, 33), (16, 16, 40)), ((5, 130, 7), (4, 4, 40)),
                    ((130, 3, 1), (2.5, 1.5, 3)), ((1, 1, 9), (1, 2, 3)), ((33, 1, 1), (3, 2, 1))]:
    lab = rng.integers(0, 4, size=shape).astype(dtype)
    # blobs rather than salt and pepper
    lab = np.repeat(np.repeat(lab[::3, ::2, :], 3, axis=0), 2, axis=1)[:shape[0], :shape[1], :]
    lab = np.asfortranarray(lab)
    _run(gpu, orc, lab, an, bb)


def test_solid_and_empty(gpu, orc):
  ones = np.ones((40, 50, 60), np.uint8, order="F")
  out, _ = _run(gpu, orc, ones, (1, 1, 1), True)
  assert out.max() == 20.0
  out, _ = _run(gpu, orc, ones, (1, 1, 1), False)
  assert np.isinf(out).all()
  _run(gpu, orc, np.zeros((17, 9, 4), np.uint32, order="F"), (1, 2, 3), True)


def test_two_missing_corners(gpu, orc):
  # the reference's test_square geometry (automated_test.py:48-63): only two background voxels
  lab = np.ones((200, 200, 1), np.uint8, order="F")
  lab[-1, 0] = 0
  lab[0, -1] = 0
  _run(gpu, orc, lab, (1, 1, 1), False)


def test_2d_border_plane(gpu, orc):
  # automated_test.py:104-114: 257^2 plane, black border, maximum at the centre
  lab = np.zeros((257, 257), np.uint32, order="F")
  lab[1:-1, 1:-1] = 1
  from kimimaro_b200 import ops
  out = ops.to_host_f(ops.edt(ops.to_device_f(lab, gpu), lab.shape, (100, 100), True), lab.shape)
  ref = orc.edt(lab, (100, 100), True)
  np.testing.assert_allclose(out, ref, rtol=RTOL)
  assert np.unravel_index(np.argmax(out), out.shape) == (128, 128)


def test_tubes_512x512x100(gpu, orc):
  # config[1]-shaped volume: full size, oracle still finishes in seconds
  from tests.synth import synthetic_tubes
  lab = synthetic_tubes((512, 512, 100), 333, seed=0xB2000333)
  _run(gpu, orc, lab, (16, 16, 40), False)


def test_bit_identical_to_oracle(gpu, orc):
  """The v2 kernels mirror the oracle's float operations one for one (same intersection formula, IEEE
  division, run-relative indices), so with exactly representable anisotropies K1 is not merely within 1e-4
  of the CPU restatement but bit-identical to it."""
  from kimimaro_b200 import ops
  from tests.synth import synthetic_tubes
  lab = synthetic_tubes((256, 192, 96), 60, seed=77)
  for an, bb in (((16, 16, 40), False), ((4, 4, 40), True), ((1, 1, 1), False)):
    out = ops.to_host_f(ops.edt(ops.to_device_f(lab, gpu), lab.shape, an, bb), lab.shape)
    ref = orc.edt(lab, an, bb)
    assert np.array_equal(out, ref)
