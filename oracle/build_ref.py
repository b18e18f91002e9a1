#!/usr/bin/env python
"""
TEST INFRASTRUCTURE -- not product code.

Compile the reference's own in-tree native extension (ext/skeletontricks) from the
sources WHERE THEY LIE under /root/reference into oracle/_ref/ (git-ignored; travels
to the GPU box with the snapshot).  Nothing is copied into the repo: cython writes its
generated .cpp into a temp dir, g++ reads the .hpp headers through -I pointing at the
reference tree, and only the resulting shared object lands in oracle/_ref/.

The extension gives a REAL oracle for the invalidation/target half of the hot path:
  roll_invalidation_ball_inside_component   (skeletontricks.pyx:373-418 ->
                                             dijkstra_invalidation.hpp:239-332)
  CachedTargetFinder                        (skeletontricks.pyx:995-1045)
  zero2inf / inf2zero / first_label         (skeletontricks.pyx:177-224, 307-326)
  find_border_targets / get_mapping         (skeletontricks.pyx:490-525, 591-715)

The reference's recipe is setup.py:12-33 (one Cython ext, -std=c++17 -O3); we do not run
its build system, we issue the two commands it amounts to.

The rest of the reference (edt, dijkstra3d, fill_voids, cc3d, osteoid ...) is NOT in
/root/reference (un-vendored PyPI deps, requirements.txt:2-7,11) and cannot be built.
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("KIMIMARO_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def ext_path():
  suffix = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
  return os.path.join(OUT, "skeletontricks" + suffix)


def build(force=False, verbose=True):
  """Returns path of the built extension, or None when the reference tree is absent."""
  target = ext_path()
  if os.path.exists(target) and not force:
    return target
  pyx = os.path.join(REF, "ext", "skeletontricks", "skeletontricks.pyx")
  if not os.path.exists(pyx):
    return None
  import numpy as np
  os.makedirs(OUT, exist_ok=True)
  with tempfile.TemporaryDirectory(prefix="b2t_ref_") as tmp:
    cpp = os.path.join(tmp, "skeletontricks.cpp")
    cmd = [sys.executable, "-m", "cython", "--cplus", "-3", pyx, "-o", cpp]
    if verbose:
      print(" ".join(cmd))
    subprocess.check_call(cmd)
    cmd = [
      "g++", "-std=c++17", "-O3", "-shared", "-fPIC", "-w",
      "-I", os.path.join(REF, "ext", "skeletontricks"),
      "-I", np.get_include(),
      "-I", sysconfig.get_paths()["include"],
      cpp, "-o", target,
    ]
    if verbose:
      print(" ".join(cmd))
    subprocess.check_call(cmd)
  return target


def load():
  """Import the compiled reference extension (None if it was never built)."""
  path = ext_path()
  if not os.path.exists(path):
    return None
  import importlib.util
  # the module's init symbol is PyInit_skeletontricks
  spec = importlib.util.spec_from_file_location("skeletontricks", path)
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod


if __name__ == "__main__":
  p = build(force="--force" in sys.argv)
  print("built:", p)
