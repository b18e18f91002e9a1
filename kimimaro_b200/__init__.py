"""
kimimaro_b200 -- B200-native (sm_100a) TEASAR skeletonization behind kimimaro's API.

  import kimimaro_b200 as kimimaro
  skels = kimimaro.skeletonize(labels, teasar_params=kimimaro.DEFAULT_TEASAR_PARAMS, anisotropy=(16, 16, 40))

mirrors kimimaro/__init__.py:18-25 for the hot path: skeletonize, synapses_to_targets, DimensionError, DEFAULT_TEASAR_PARAMS
and the Skeleton result type.  Everything numeric runs in hand-written CUDA kernels reached through
the C ABI of libb2t.so (include/b2t.h); importing this package does not need a GPU, calling it does.
"""
from .skeleton import Skeleton  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
  # lazy: `import kimimaro_b200` must work on a CPU-only box (build / CI), torch is imported on first use
  if name in ("skeletonize", "DimensionError", "DEFAULT_TEASAR_PARAMS", "synapses_to_targets", "connect_points"):
    from . import intake
    return getattr(intake, name)
  if name in ("set_invalidation_mode", "invalidation_mode"):    # "window" (default) | "strict" (the reference's heap order)
    from . import _lib
    return getattr(_lib, name)
  if name in ("postprocess", "join_close_components"):            # chunk-stitch post-processing (kimimaro/__init__.py:23)
    from . import post
    return getattr(post, name)
  if name == "post":
    import importlib
    return importlib.import_module(".post", __name__)
  if name == "skeletonize_sharded":                             # one process per GPU, labels sharded over the ranks
    from . import distributed
    return distributed.skeletonize_sharded
  raise AttributeError(name)
