"""Round-2 check of the key-ordered invalidation rounds (trace.cu: invalidate_window, built only into the
claim_window variant of the library):

  python -m kimimaro_b200.build --variant claim_window
  B2T_LIB=$PWD/kimimaro_b200/_variants/claim_window.so B2T_CLAIM_WINDOW=1 python scripts/claim_window_parity.py

Tier A: the CUDA path with the window on must equal the oracle's mode "window:1" bit for bit (vertices, edges).
Tier B: how many skeletons equal the reference's literal heap order (oracle mode "heap" == the compiled reference
extension voxel for voxel) -- on the CPU the hop rounds reach 126 of 176, the key rounds 174 of 176 -- and what the
path loop costs with either."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import kimimaro_b200
from kimimaro_b200 import _lib
from kimimaro_b200.datasets import synthetic_tubes
from oracle import teasar


def vertex_sets(res, get):
  return {k: {tuple(x) for x in get(v).tolist()} for k, v in res.items()}


def main():
  window = float(os.environ.get("B2T_CLAIM_WINDOW", "1"))
  _lib.require_device()
  _lib.check(_lib.lib().b2t_set_claim_window(_lib.c_f32(window)), "b2t_set_claim_window")
  out = []
  for seed, shape, n in ((1, (192, 160, 96), 60), (2, (256, 192, 96), 80), (3, (160, 160, 160), 70)):
    lab = synthetic_tubes(shape, n, seed=seed)
    tm = {}
    torch.cuda.synchronize()
    t = time.time()
    got = kimimaro_b200.skeletonize(lab, anisotropy=(16, 16, 40), progress=False, timings=tm)
    torch.cuda.synchronize()
    dt = time.time() - t
    same_mode = teasar.skeletonize(lab, anisotropy=(16, 16, 40), invalidation_mode=f"window:{window:g}", parallel=8)
    heap = teasar.skeletonize(lab, anisotropy=(16, 16, 40), invalidation_mode="heap", parallel=8)
    tier_a = sorted(got) == sorted(same_mode) and all(
      np.array_equal(got[k].vertices, same_mode[k]["vertices"]) and np.array_equal(got[k].edges, same_mode[k]["edges"])
      for k in same_mode)
    g, h = vertex_sets(got, lambda s: s.vertices), vertex_sets(heap, lambda s: s["vertices"])
    rec = {"seed": seed, "skeletons": len(heap), "tier_a_bit_exact_vs_oracle_window": bool(tier_a),
           "identical_to_reference_heap_order": int(sum(g.get(k) == h[k] for k in h)), "seconds": dt,
           "path_loop_s": tm.get("paths")}
    print(json.dumps(rec), flush=True)
    out.append(rec)
  os.makedirs("gpurun_out", exist_ok=True)
  with open("gpurun_out/claim_window_parity.json", "w") as f:
    json.dump(out, f, indent=1)


main()
