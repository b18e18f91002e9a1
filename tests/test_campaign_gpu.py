"""A short run of the randomised parity campaign (scripts/gpu_parity_campaign.py): random shapes, anisotropies (one
non-integer), label dtypes and options; the default order against the oracle in the same order and the strict mode
against the reference's heap order.  The long runs are under profiles/ (r02_gpu_parity_campaign_*.jsonl)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_random_cases_default_and_strict(gpu):
  p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gpu_parity_campaign.py"), "8", "7"], capture_output=True,
                     text=True, timeout=900, cwd=ROOT)
  assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
  s = json.loads(p.stdout.strip().splitlines()[-1])
  assert s["summary"] and s["default_equals_oracle"] == 8 and s["strict_equals_reference_heap_order"] == 8
