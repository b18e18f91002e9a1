"""Where a label's time goes inside the path loop: per-phase SM cycles of the 'prof' variant build
(python -m kimimaro_b200.build --variant prof), largest labels of the benchmark volume.
  B2T_LIB=kimimaro_b200/_variants/prof.so python scripts/trace_prof.py [size=512]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import kimimaro_b200
from kimimaro_b200 import _lib
from bench import make_volume, anisotropy_of

NAMES = ["find_target", "rr_min", "rr_split", "rr_expand", "rr_round_end", "rr_refill", "rr_walk", "rr_reset",
         "inv_push", "inv_kmin", "inv_claim", "misc", "n_inv_rounds", "n_refills", "-", "-"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
vol = make_volume(n)
an = anisotropy_of(n)
if os.environ.get("B2T_AN"):                 # e.g. B2T_AN=4,4,40: the blob of the volume becomes a giant ordinary label
  an = tuple(float(v) for v in os.environ["B2T_AN"].split(","))
kimimaro_b200.skeletonize(vol, anisotropy=an, progress=False)
tm = {}
kimimaro_b200.skeletonize(vol, anisotropy=an, progress=False, timings=tm)
torch.cuda.synchronize()
buf = np.zeros((64, 16), dtype=np.uint64)
fn = _lib.lib().b2t_trace_prof_read
fn.argtypes = [ctypes.c_void_p]
_lib.check(fn(buf.ctypes.data), "b2t_trace_prof_read")
mhz = 1965.0
if len(tm["kernel_stats"]) > 1:              # the table's row 0 now belongs to the LAST batch: the private arena's one label
  sp = tm["kernel_stats"][-1]
  rec = {"job": "private", "us": int(sp["stats"][0, 3]), "npaths": int(sp["npaths"][0]), "rr_rounds": int(sp["stats"][0, 1]),
         "relax": int(sp["stats"][0, 0]), "invalidated": int(sp["stats"][0, 2])}
  for k, name in enumerate(NAMES):
    if name != "-":
      rec[name] = int(buf[0][k]) if name.startswith("n_") else round(float(buf[0][k]) / mhz, 1)
  rec["phases_ms"] = {k: round(1e3 * v, 2) for k, v in tm.items() if isinstance(v, float) and k.startswith("soma")}
  print(json.dumps(rec), flush=True)
st = tm["kernel_stats"][0]
for job in range(1, 24):
  row = buf[job]
  rec = {"job": job, "us": int(st["stats"][job, 3]), "npaths": int(st["npaths"][job]), "rr_rounds": int(st["stats"][job, 1]),
         "relax": int(st["stats"][job, 0]), "invalidated": int(st["stats"][job, 2])}
  for k, name in enumerate(NAMES):
    if name == "-":
      continue
    rec[name] = int(row[k]) if name.startswith("n_") else round(float(row[k]) / mhz, 1)     # microseconds
  print(json.dumps(rec), flush=True)
