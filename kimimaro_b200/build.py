"""
Build kimimaro_b200/libb2t.so (hand-written CUDA for sm_100a + the C ABI of include/b2t.h).

  python -m kimimaro_b200.build [--force] [--verbose]

In-tree, explicit nvcc: the .so sits next to this file (git-ignored, but it travels with the
gpurun snapshot).  No JIT cache, no torch.utils.cpp_extension, no multi-arch fatbin.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb2t.so")
SOURCES = ["capi.cu", "edt.cu", "field.cu", "trace.cu", "preamble.cu"]
HEADERS = ["common.cuh", "edt_fh3.cuh", os.path.join("..", "..", "include", "b2t.h")]

NVCC_FLAGS = [
  "-gencode", "arch=compute_100a,code=sm_100a",
  "-O3", "-lineinfo", "-std=c++17",
  "-fmad=false",              # float32 expressions must round exactly like the reference's scalar code
  "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
  "-shared", "-cudart", "shared",
]


def sources():
  return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _stale():
  if not os.path.exists(LIB):
    return True
  t = os.path.getmtime(LIB)
  deps = sources() + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
  return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


# other builds of the same sources for kernel A/B runs (B2T_LIB=kimimaro_b200/_variants/<name>.so selects one)
VARIANTS = {
  "claim_window": ["-DB2T_WITH_CLAIM_WINDOW=1"],   # key-ordered invalidation rounds in the path loop (trace.cu)
}


def build_variant(name, verbose=False):
  out_dir = os.path.join(HERE, "_variants")
  os.makedirs(out_dir, exist_ok=True)
  out = os.path.join(out_dir, name + ".so")
  nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
  cmd = [nvcc] + NVCC_FLAGS + VARIANTS[name] + (["-Xptxas", "-v"] if verbose else []) + sources() + ["-o", out]
  if verbose:
    print(" ".join(cmd))
  subprocess.check_call(cmd)
  return out


def build(force=False, verbose=False):
  if not force and not _stale():
    return LIB
  nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
  cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + sources() + ["-o", LIB]
  if verbose:
    print(" ".join(cmd))
  subprocess.check_call(cmd)
  return LIB


if __name__ == "__main__":
  if "--variant" in sys.argv:
    print(build_variant(sys.argv[sys.argv.index("--variant") + 1], verbose="--verbose" in sys.argv))
  else:
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
