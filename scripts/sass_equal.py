"""Is the DEVICE code of the working tree the device code of an earlier commit?  Compiles every kimimaro_b200/csrc/*.cu of
both for sm_100a (no GPU needed), disassembles with cuobjdump and compares kernel by kernel (names normalised for the
anonymous-namespace hash and for template parameters that were appended with a default since).  Used at the end of
round 1, when the GPU budget was spent, to show that the shipped kernels are byte for byte the ones the last GPU run
tested:

  python scripts/sass_equal.py 0b68cae        # the build of profiles/r01_k1_shot.jsonl and r01_roles_gpu_tests.log
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false"]


def sass(cu, out_dir, tag):
  obj = os.path.join(out_dir, tag + ".o")
  subprocess.check_call(["nvcc"] + FLAGS + ["-c", cu, "-o", obj])
  txt = subprocess.check_output(["cuobjdump", "-sass", obj]).decode()
  funcs, cur = {}, None
  for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
      cur = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_[a-z_0-9]+_cu_[0-9a-f]+", "NS", m.group(1))
      cur = re.sub(r"(fh3_range_kernelIjLi\d+ELi\d+ELi\d+ELi\d+ELi\d+E)Li1ELi0E", r"\1", cur)   # QP = 1, PP = 0 appended since
      funcs[cur] = []
    elif cur and line.strip() and not line.strip().startswith("."):
      funcs[cur].append(re.sub(r"/\*[0-9a-f]{4,}\*/", "", line).strip())
  return funcs


def main():
  commit = sys.argv[1]
  bad = 0
  with tempfile.TemporaryDirectory() as tmp:
    old = os.path.join(tmp, "old", "kimimaro_b200", "csrc")
    os.makedirs(old)
    os.makedirs(os.path.join(tmp, "old", "include"))
    for rel in ["include/b2t.h"] + ["kimimaro_b200/csrc/" + f for f in os.listdir(os.path.join(ROOT, "kimimaro_b200", "csrc"))]:
      try:
        data = subprocess.check_output(["git", "-C", ROOT, "show", f"{commit}:{rel}"], stderr=subprocess.DEVNULL)
      except subprocess.CalledProcessError:
        continue
      with open(os.path.join(tmp, "old", rel), "wb") as f:
        f.write(data)
    for name in sorted(os.listdir(os.path.join(ROOT, "kimimaro_b200", "csrc"))):
      if not name.endswith(".cu") or not os.path.exists(os.path.join(old, name)):
        continue
      a = sass(os.path.join(old, name), tmp, "old_" + name)
      b = sass(os.path.join(ROOT, "kimimaro_b200", "csrc", name), tmp, "new_" + name)
      same = [k for k in a if k in b and a[k] == b[k]]
      diff = [k for k in a if k in b and a[k] != b[k]]
      gone = [k for k in a if k not in b]
      print(f"{name}: {len(a)} kernels at {commit}, {len(b)} now; identical {len(same)}, changed {len(diff)}, "
            f"gone {len(gone)}, new {len(b) - len(same) - len(diff)}")
      for k in diff + gone:
        print("   ", k[:140])
      bad += len(diff) + len(gone)
  sys.exit(1 if bad else 0)


main()
