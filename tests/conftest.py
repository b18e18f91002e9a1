import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
  """The CPU oracle (test infrastructure)."""
  import oracle
  oracle.build()
  return oracle


@pytest.fixture(scope="session")
def ref_ext():
  """The reference's own in-tree extension compiled into oracle/_ref (None if never built)."""
  from oracle import build_ref
  try:
    build_ref.build(verbose=False)
  except Exception:
    pass
  return build_ref.load()


@pytest.fixture(scope="session")
def gpu():
  import torch
  from kimimaro_b200 import _lib
  _lib.require_device()
  return torch.device("cuda:0")


def golden_name(stem, ext):
  """tests/golden/<stem>[_<mode>]<ext> for the claim order the oracle (and so the engine under test) runs by default."""
  import os
  from oracle import teasar
  mode = teasar.DEFAULT_INVALIDATION_MODE
  suffix = "" if mode == "rounds" else "_" + mode.replace(":", "").replace(".", "p")
  return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", stem + suffix + ext)
