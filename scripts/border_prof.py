"""cProfile of engine.compute_border_targets on the benchmark volume (host-side cost of the six faces)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from kimimaro_b200 import engine
from bench import make_volume, anisotropy_of

vol = make_volume(512)
an = anisotropy_of(512)
d = torch.from_numpy(vol.reshape(-1, order="F").view(np.int32)).cuda()
cc, n_cc = engine.connected_components(d, vol.shape)
for _ in range(3):
  engine.compute_border_targets(cc, vol.shape, an)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(10):
  engine.compute_border_targets(cc, vol.shape, an)
torch.cuda.synchronize()
print("ms per call:", 100 * (time.perf_counter() - t))
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
  engine.compute_border_targets(cc, vol.shape, an)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
