"""
Device-side operators of the TEASAR hot path, one Python function per native call the reference
makes in kimimaro/trace.py / kimimaro/intake.py, each a thin wrapper over one C-ABI entry point of
libb2t.so.  torch tensors are used only as device-memory containers (data_ptr()).

Layout: every volume tensor is the FLAT Fortran-order buffer of a [sx,sy,sz] array,
loc = x + sx*(y + sy*z) (kimimaro/intake.py:320-322).
"""
import numpy as np
import torch

from ._lib import c_f32, c_i64, c_int, c_sz, c_vp, check, lib, stream_ptr

def _ptr(t):
  return c_vp(t.data_ptr())


def to_device_f(arr, device="cuda"):
  """numpy [sx,sy,sz] (any order) -> flat Fortran-order device tensor (one H2D copy)."""
  a = np.asfortranarray(arr)
  flat = a.reshape(-1, order="F")
  if flat.dtype == bool:
    flat = flat.view(np.uint8)
  return torch.from_numpy(flat).to(device, non_blocking=True)


def to_host_f(t, shape):
  """flat Fortran-order device tensor -> numpy [shape] Fortran-ordered."""
  return t.cpu().numpy().reshape(shape, order="F")


def edt(d_labels, shape, anisotropy=(1.0, 1.0, 1.0), black_border=False, out=None, workspace=True):
  """edt.edt(labels, anisotropy=, black_border=) -- kimimaro/intake.py:174-185, trace.py:112-117.
  d_labels: flat device tensor of an unsigned integer dtype; shape: (sx,sy) or (sx,sy,sz).
  2-D shapes run the 2-D transform like the library (intake.py:565)."""
  ndim = len(shape)
  if ndim not in (2, 3):
    raise ValueError(f"edt: shape must have 2 or 3 extents, got {tuple(shape)}")
  sx, sy = int(shape[0]), int(shape[1])
  sz = int(shape[2]) if ndim == 3 else 1
  an = [float(a) for a in anisotropy] + [1.0] * (3 - len(anisotropy))
  if not (d_labels.is_cuda and d_labels.is_contiguous() and d_labels.numel() == sx * sy * sz):
    raise ValueError("edt: d_labels must be a contiguous device tensor of sx * sy * sz elements")
  if out is None:
    out = torch.empty(sx * sy * sz, dtype=torch.float32, device=d_labels.device)
  # uint32 labels get a scratch volume so that the column passes can run as stencil + envelope (b2t_edt_ws);
  # the caching allocator makes this a stream-ordered pointer bump, and the library falls back to the in-place
  # path by itself when the input does not qualify
  if workspace and d_labels.element_size() == 4:
    nbytes = int(lib().b2t_edt_workspace_bytes(c_i64(sx), c_i64(sy), c_i64(sz)))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=d_labels.device)
    check(lib().b2t_edt_ws(_ptr(d_labels), c_int(4), c_i64(sx), c_i64(sy), c_i64(sz), c_f32(an[0]), c_f32(an[1]),
                           c_f32(an[2]), c_int(int(bool(black_border))), c_int(ndim), _ptr(out), _ptr(ws),
                           c_sz(nbytes), stream_ptr()), "b2t_edt_ws")
    return out
  check(lib().b2t_edt(_ptr(d_labels), c_int(d_labels.element_size()), c_i64(sx), c_i64(sy), c_i64(sz),
                      c_f32(an[0]), c_f32(an[1]), c_f32(an[2]), c_int(int(bool(black_border))), c_int(ndim),
                      _ptr(out), stream_ptr()), "b2t_edt")
  return out
