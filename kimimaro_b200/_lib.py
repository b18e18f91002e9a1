"""
ctypes binding of libb2t.so (include/b2t.h).  Fails loudly: no library or no sm_100 device means
an exception, never a CPU fallback.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# B2T_LIB: another build of the same sources (kernel experiments); never a different implementation
LIB_PATH = os.environ.get("B2T_LIB") or os.path.join(HERE, "libb2t.so")

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_f32 = ctypes.c_float
c_vp = ctypes.c_void_p
c_sz = ctypes.c_size_t

_lib = None


class B2TError(RuntimeError):
  pass


# name -> argtypes; every function returns int status unless listed in _RESTYPES
_SIGNATURES = {
  "b2t_version": [],
  "b2t_last_error": [],
  "b2t_device_check": [],
  "b2t_edt": [c_vp, c_int, c_i64, c_i64, c_i64, c_f32, c_f32, c_f32, c_int, c_int, c_vp, c_vp],
  "b2t_edt_config": [c_int, c_int, c_int, c_int, c_int],
  "b2t_edt_config_hybrid": [c_int, c_int, c_int, c_int, c_int, c_int],
  "b2t_edt_config_roles": [c_int, c_int, c_f32],
  "b2t_edt_config_envelope": [c_int, c_int],
  "b2t_edt_workspace_bytes": [c_i64, c_i64, c_i64],
  "b2t_edt_ws": [c_vp, c_int, c_i64, c_i64, c_i64, c_f32, c_f32, c_f32, c_int, c_int, c_vp, c_vp, c_sz, c_vp],
}
_RESTYPES = {"b2t_last_error": ctypes.c_char_p, "b2t_edt_workspace_bytes": c_sz}


def declare(name, argtypes, restype=None):
  """Used by the op modules to register further entry points of include/b2t.h."""
  _SIGNATURES[name] = argtypes
  if restype is not None:
    _RESTYPES[name] = restype
  if _lib is not None:
    fn = getattr(_lib, name)
    fn.argtypes = argtypes
    fn.restype = _RESTYPES.get(name, c_int)


def lib():
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise B2TError(
        f"{LIB_PATH} is missing: build it with `python -m kimimaro_b200.build` "
        "(kimimaro_b200 has no CPU fallback)")
    _lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
      fn = getattr(_lib, name)
      fn.argtypes = argtypes
      fn.restype = _RESTYPES.get(name, c_int)
  return _lib


def check(status, what=""):
  if status != 0:
    msg = lib().b2t_last_error()
    raise B2TError(f"{what} failed with status {status}: {msg.decode() if msg else ''}")


_checked_devices = set()


def require_device():
  """Raise unless the current device is an sm_100 GPU.  The answer is cached per device:
  cudaGetDeviceProperties is slow and jittery (tens of ms, occasionally hundreds) and must not sit on the
  per-call path of skeletonize()."""
  import torch
  if not torch.cuda.is_available():
    raise B2TError("kimimaro_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
  dev = torch.cuda.current_device()
  if dev in _checked_devices:
    return
  check(lib().b2t_device_check(), "b2t_device_check")
  _checked_devices.add(dev)


# ---- claim order of roll_invalidation_ball_inside_component in the path loop (include/b2t.h: B2T_INVALIDATE_*) ----
INVALIDATE_ROUNDS, INVALIDATE_WINDOW, INVALIDATE_STRICT = 0, 1, 2
_MODES = {"rounds": INVALIDATE_ROUNDS, "window": INVALIDATE_WINDOW, "strict": INVALIDATE_STRICT}
_invalidation = None


def set_invalidation_mode(mode, window=1.0):
  """'window' (default): parallel rounds ordered by the reference's heap key in windows of `window` voxel edges;
  'strict': the reference's priority queue literally (identical to the compiled reference, sequential per label);
  'rounds': hop-synchronous rounds.  The environment variable B2T_INVALIDATION ('strict', 'rounds', 'window[:w]')
  sets the initial value."""
  global _invalidation
  if mode not in _MODES:
    raise ValueError(f"invalidation mode must be one of {sorted(_MODES)}, got {mode!r}")
  if mode == "window" and not window > 0:
    raise ValueError("the window mode needs a positive window")
  _invalidation = (mode, float(window))


def invalidation_mode():
  """(mode name, window in voxel edges)"""
  if _invalidation is None:
    env = os.environ.get("B2T_INVALIDATION", "window").split(":")
    set_invalidation_mode(env[0], float(env[1]) if len(env) > 1 else 1.0)
  return _invalidation


_raw_stream = None


def stream_ptr():
  """cudaStream_t of torch's current stream on the current device (every C call takes it).  torch.cuda.current_stream()
  resolves the device through several Python layers (~15 us, a hundred times per pass); the raw getter does not."""
  global _raw_stream
  import torch
  if _raw_stream is None:
    getter = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    _raw_stream = getter if getter is not None else False
  if _raw_stream:
    try:
      return c_vp(_raw_stream(torch.cuda.current_device()))
    except Exception:             # no CUDA runtime behind torch (the CPU suite's emulated library stubs current_stream)
      _raw_stream = False
  return c_vp(torch.cuda.current_stream().cuda_stream)
