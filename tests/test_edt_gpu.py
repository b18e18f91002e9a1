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


# ---- b2t_edt_ws: stencil + envelope hybrid (uint32 labels, integer anisotropy) and its fallbacks ----
def test_hybrid_bit_identical_to_oracle(gpu, orc):
  """ops.edt hands uint32 volumes to b2t_edt_ws; with integer anisotropy the column passes run as the register
  stencil plus the envelope kernel on the flagged blocks, and the result is the oracle's, bit for bit."""
  from kimimaro_b200 import ops
  from tests.synth import synthetic_tubes
  lab = synthetic_tubes((256, 192, 96), 60, seed=78).astype(np.uint32)
  lab[60:200, 40:150, 20:80] = 4242                       # a blob: rows the stencil cannot finish
  lab = np.asfortranarray(lab)
  for an, bb in (((16, 16, 40), False), ((4, 4, 40), True), ((1, 1, 1), False), ((40, 32, 20), True), ((2, 3, 5), False)):
    out = ops.to_host_f(ops.edt(ops.to_device_f(lab, gpu), lab.shape, an, bb, workspace=True), lab.shape)
    assert np.array_equal(out, orc.edt(lab, an, bb)), (an, bb)
  # dense labels (a run is always open), ragged sizes, columns shorter than the stencil window, a 2-D plane
  rng = np.random.default_rng(4)
  for shape in ((64, 70, 33), (128, 5, 3), (32, 300, 2), (4, 3, 200), (96, 1, 1)):
    dense = rng.integers(1, 6, size=shape).astype(np.uint32)
    dense = np.asfortranarray(np.repeat(np.repeat(dense[:, ::4, ::3], 4, axis=1), 3, axis=2)[:, :shape[1], :shape[2]])
    for bb in (False, True):
      out = ops.to_host_f(ops.edt(ops.to_device_f(dense, gpu), dense.shape, (16, 16, 40), bb, workspace=True), dense.shape)
      assert np.array_equal(out, orc.edt(dense, (16, 16, 40), bb)), (shape, bb)
  plane = np.zeros((260, 257), np.uint32, order="F")
  plane[1:-1, 1:-1] = 1
  plane[100:140, 50:200] = 2
  out = ops.to_host_f(ops.edt(ops.to_device_f(plane, gpu), plane.shape, (100, 100), True, workspace=True), plane.shape)
  assert np.array_equal(out, orc.edt(plane, (100, 100), True))


def test_hybrid_equals_in_place_path(gpu):
  """Same input through b2t_edt_ws and b2t_edt: identical floats (the hybrid is an execution strategy, not an
  approximation); also with the hybrid switched off and with a workspace that is too small (both fall back)."""
  import torch
  from kimimaro_b200 import ops, _lib
  from kimimaro_b200._lib import c_f32, c_i64, c_int, c_sz, c_vp
  from tests.synth import synthetic_tubes
  lab = np.asfortranarray(synthetic_tubes((192, 160, 80), 40, seed=79).astype(np.uint32))
  lab[30:150, 30:130, 10:70] = 99
  d = ops.to_device_f(lab, gpu)
  ref = ops.edt(d, lab.shape, (16, 16, 40), False, workspace=False).clone()
  assert torch.equal(ops.edt(d, lab.shape, (16, 16, 40), False, workspace=True), ref)
  lib = _lib.lib()
  try:
    _lib.check(lib.b2t_edt_config_hybrid(0, -1, -1, 0, 0, 0))
    assert torch.equal(ops.edt(d, lab.shape, (16, 16, 40), False, workspace=True), ref)
  finally:
    _lib.check(lib.b2t_edt_config_hybrid(1, -1, -1, 0, 0, 0))
  out = torch.empty_like(ref)
  ws = torch.empty(1024, dtype=torch.uint8, device=gpu)
  _lib.check(lib.b2t_edt_ws(c_vp(d.data_ptr()), c_int(4), c_i64(192), c_i64(160), c_i64(80), c_f32(16), c_f32(16), c_f32(40),
                            c_int(0), c_int(3), c_vp(out.data_ptr()), c_vp(ws.data_ptr()), c_sz(1024), ops.stream_ptr()),
             "b2t_edt_ws")
  assert torch.equal(out, ref)


def test_k1_forms_identical(gpu):
  """The execution forms of b2t_edt_ws selectable through b2t_edt_config_roles -- both stencil bodies, with and
  without the roles launch (prediction from the previous pass, envelope and stencil warps side by side) -- give the
  same floats as the shipped default; a prediction scaled to miss blocks only moves work to the residual launch."""
  import torch
  from kimimaro_b200 import ops, _lib
  from kimimaro_b200._lib import c_f32
  from tests.synth import synthetic_tubes
  lib = _lib.lib()
  lab = np.asfortranarray(synthetic_tubes((192, 160, 96), 40, seed=80).astype(np.uint32))
  lab[30:150, 30:130, 10:80] = 99
  plane = np.zeros((260, 257), np.uint32, order="F")
  plane[1:-1, 1:-1] = 1
  plane[100:140, 50:200] = 2
  cases = [(lab, (16, 16, 40), False), (lab, (4, 4, 40), True), (lab, (1, 1, 1), False), (lab, (40, 32, 20), True),
           (plane, (100, 100), True)]
  devs = [ops.to_device_f(a, gpu) for a, _, _ in cases]
  refs = [ops.edt(d, a.shape, an, bb, workspace=True).clone() for d, (a, an, bb) in zip(devs, cases)]
  try:
    for roles, v2, scale in ((0, 0, 1.0), (1, 0, 1.0), (1, 1, 1.0), (1, 1, 3.0)):
      _lib.check(lib.b2t_edt_config_roles(roles, v2, c_f32(scale)))
      for d, (a, an, bb), ref in zip(devs, cases, refs):
        assert torch.equal(ops.edt(d, a.shape, an, bb, workspace=True), ref), (roles, v2, scale, an, bb)
  finally:
    _lib.check(lib.b2t_edt_config_roles(0, 1, c_f32(1.0)))   # the shipped default


@pytest.mark.parametrize("dtype", [np.uint8, np.uint32])
def test_hybrid_fallbacks_keep_tolerance(gpu, orc, dtype):
  # non-integer anisotropy, narrow labels, sx not a multiple of 4: b2t_edt_ws takes the b2t_edt path
  from tests.synth import synthetic_tubes
  for shape, an in (((96, 80, 48), (3.3, 4.7, 10.1)), ((97, 64, 40), (16, 16, 40))):
    lab = np.asfortranarray((synthetic_tubes(shape, 12, seed=6) % 200).astype(dtype))
    _run(gpu, orc, lab, an, False)


def test_x_pass_tma_form_is_bit_identical(gpu):
  """b2t_edt_config_xpass(1): the x pass of b2t_edt_ws stages its tiles with TMA (edt_xtma.cuh: cp.async.bulk.tensor loads
  and stores behind an mbarrier).  Same bits as the register-only kernel on rows of 512 and of 256 labels, with and
  without a black border, and the launch counter proves that the TMA kernel is the one that ran."""
  import ctypes
  import torch
  from kimimaro_b200 import ops, _lib
  from tests.synth import synthetic_tubes
  L = _lib.lib()
  L.b2t_edt_config_xpass.restype = ctypes.c_longlong
  L.b2t_edt_config_xpass.argtypes = [ctypes.c_int]
  try:
    for shape, an, bb in (((512, 96, 40), (16, 16, 40), False), ((256, 64, 48), (4, 4, 40), True), ((512, 33, 7), (1, 1, 1), False)):
      lab = synthetic_tubes(shape, 12, seed=5).astype(np.uint32)
      d = ops.to_device_f(lab)
      L.b2t_edt_config_xpass(0)
      a = ops.edt(d, shape, an, bb).clone()
      for form in (1, 2):                                    # a tile per CTA; persistent CTAs with a two-stage ring
        before = L.b2t_edt_config_xpass(form)
        b = ops.edt(d, shape, an, bb).clone()
        torch.cuda.synchronize()
        assert L.b2t_edt_config_xpass(-1) == before + 1, "the TMA x pass did not launch"
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), form
  finally:
    L.b2t_edt_config_xpass(0)
