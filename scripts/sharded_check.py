"""On-hardware multi-GPU parity (the analogue of the reference's test_parallel, automated_test.py:234-259):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/sharded_check.py
Every rank traces its LPT share of the connected components, rank 0 gathers the path buffers over NCCL, assembles the skeletons and
compares the merged result with what it computes alone on the same volume: identical ids, vertices, edges, radii."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import kimimaro_b200
from kimimaro_b200 import distributed as kd
from kimimaro_b200.datasets import synthetic_tubes


def main():
  rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  dist.init_process_group("nccl", device_id=dev)
  lab = synthetic_tubes((256, 256, 128), 120, seed=31, anisotropy=(16, 16, 40))
  kw = dict(anisotropy=(16, 16, 40), progress=False, dust_threshold=300)
  tm = {}
  out = kd.skeletonize_sharded(lab, timings=tm, **kw)        # slab upload + all-gather, LPT share, one gather of path buffers
  mine = range(tm.get("n_traced", 0))
  ok = True
  if rank == 0:
    alone = kimimaro_b200.skeletonize(lab, **kw)
    ok = sorted(out) == sorted(alone) and len(mine) < len(alone) and all(
      np.array_equal(out[k].vertices, alone[k].vertices) and np.array_equal(out[k].edges, alone[k].edges)
      and np.array_equal(out[k].radii, alone[k].radii) for k in alone)
    print(json.dumps({"sharded_equals_single": bool(ok), "world": world, "skeletons": len(alone),
                      "traced_on_rank0": len(mine)}), flush=True)
  dist.barrier()
  dist.destroy_process_group()
  sys.exit(0 if ok else 1)


main()
