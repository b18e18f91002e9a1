"""N > 1 on real GPUs (skipped on a one-GPU box): labels sharded over NCCL ranks must give rank 0 exactly the skeletons
one GPU computes (scripts/sharded_check.py under torchrun).  The gloo / CPU twin is tests/test_distributed_cpu.py and
tests/test_product_on_emulated_library_cpu.py::test_sharded_over_two_ranks_gloo."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_equals_single_on_nccl(gpu):
  import torch
  n = torch.cuda.device_count()
  if n < 2:
    pytest.skip("needs at least two GPUs")
  n = 2 if n < 4 else 4
  from tests.test_distributed_cpu import _free_port
  cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
         "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", "sharded_check.py")]
  p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
  assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
  line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
  assert json.loads(line)["sharded_equals_single"] is True
