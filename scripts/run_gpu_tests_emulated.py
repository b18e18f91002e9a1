"""Run the bodies of GPU-marked tests on the CPU: kimimaro_b200 pointed at the emulated library
(tests/test_product_on_emulated_library_cpu.py: the library's kernels compiled for the CPU against the SIMT emulation)
with CPU tensors standing in for device memory.  Slow (an emulated 512-thread block per label), so it is a script and
not part of the CPU suite.  Used at the end of round 1 to run tests/test_zz_options_gpu.py, written after the round's
GPU budget was spent, before its first run on a device:

  python scripts/run_gpu_tests_emulated.py tests.test_zz_options_gpu [test_name ...]
"""
import importlib
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tests.test_product_on_emulated_library_cpu as T  # noqa: E402

T._emulate(setattr, T._build())
mod = importlib.import_module(sys.argv[1])
names = sys.argv[2:] or [n for n in dir(mod) if n.startswith("test_")]
failed = 0
for name in names:
  t = time.time()
  try:
    getattr(mod, name)("cpu")          # the `gpu` fixture: a device
    print(name, "PASSED", round(time.time() - t, 1), "s", flush=True)
  except Exception:
    traceback.print_exc()
    print(name, "FAILED", flush=True)
    failed += 1
sys.exit(1 if failed else 0)
