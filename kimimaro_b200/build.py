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
HEADERS = ["common.cuh", "edt_fh3.cuh", "edt_xtma.cuh", os.path.join("..", "..", "include", "b2t.h")]

NVCC_FLAGS = [
  "-gencode", "arch=compute_100a,code=sm_100a",
  "-O3", "-lineinfo", "-std=c++17",
  "-fmad=false",              # float32 expressions must round exactly like the reference's scalar code
  "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
]
LINK_FLAGS = ["-shared", "-cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a"]
OBJ_DIR = os.path.join(HERE, "build")


def sources():
  return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _stale():
  if not os.path.exists(LIB):
    return True
  t = os.path.getmtime(LIB)
  deps = sources() + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
  return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile(extra, out, verbose=False, force=False):
  """One object per source, compiled in parallel and cached by modification time (edt.cu alone takes a minute), then
  one link.  No relocatable device code: the files share no device symbols."""
  from concurrent.futures import ThreadPoolExecutor
  nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
  tag = "_".join(e.replace("-D", "").replace("=", "") for e in extra) or "default"
  odir = os.path.join(OBJ_DIR, tag)
  os.makedirs(odir, exist_ok=True)
  hdr_t = max(os.path.getmtime(d) for d in [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)])

  def one(src):
    obj = os.path.join(odir, os.path.basename(src) + ".o")
    if (not force) and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_t):
      return obj
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    if verbose:
      print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return obj

  with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
    objs = list(ex.map(one, sources()))
  subprocess.check_call([nvcc] + LINK_FLAGS + objs + ["-o", out])
  return out


# other builds of the same sources for kernel A/B runs (B2T_LIB=kimimaro_b200/_variants/<name>.so selects one)
VARIANTS = {
  "prof": ["-DB2T_TRACE_PROF"],       # per-phase cycle counters in the path loop (scripts/trace_prof.py)
  "rr_global": ["-DB2T_RR_SOLO=0"],   # solo CTAs keep railroad's near lists in global memory (the form before railroad_solo)
  "minb3": ["-DB2T_TRACE_MINB=3"],    # 42 registers per thread in the path loop (three resident CTAs per SM; the default until call 25)
  "minb2_cap1k": ["-DB2T_RR_SOLO_CAP=1024"],                         # smaller shared-memory lists: more L1 for the spills
  "t256_minb3_cap1k": ["-DB2T_TRACE_THREADS=256", "-DB2T_TRACE_MINB=3", "-DB2T_RR_SOLO_CAP=1024"],
  "batch2": ["-DB2T_RR_BATCH=2"],
  "edf_solo": ["-DB2T_EDF_SOLO=1"],   # per-label sweeps with the frontier in shared memory (slower: see field.cu)
  "pdrf_cg": ["-DB2T_PDRF_CA=0"],     # railroad_solo gathers the PDRF from L2 only (the form before call 58)
  "minb1": ["-DB2T_TRACE_MINB=1"],    # 128 registers, one resident CTA per SM
  "t256_minb3": ["-DB2T_TRACE_THREADS=256", "-DB2T_TRACE_MINB=3"],   # 85 registers
  "t256_minb2": ["-DB2T_TRACE_THREADS=256", "-DB2T_TRACE_MINB=2"],   # 128 registers
}


def build_variant(name, verbose=False):
  out_dir = os.path.join(HERE, "_variants")
  os.makedirs(out_dir, exist_ok=True)
  return _compile(VARIANTS[name], os.path.join(out_dir, name + ".so"), verbose=verbose)


def build(force=False, verbose=False):
  if not force and not _stale():
    return LIB
  return _compile([], LIB, verbose=verbose, force=force)


if __name__ == "__main__":
  if "--variant" in sys.argv:
    print(build_variant(sys.argv[sys.argv.index("--variant") + 1], verbose="--verbose" in sys.argv))
  else:
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
