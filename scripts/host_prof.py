"""cProfile of kimimaro_b200.skeletonize on the benchmark volume (labels resident): where the HOST spends the pass."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import kimimaro_b200
from bench import make_volume, anisotropy_of

vol = make_volume(512)
an = anisotropy_of(512)
d = torch.from_numpy(vol.reshape(-1, order="F").view(np.int32)).cuda()
kw = dict(device_labels=d, anisotropy=an, progress=False)
for _ in range(3):
  kimimaro_b200.skeletonize(vol.shape, **kw)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5):
  kimimaro_b200.skeletonize(vol.shape, **kw)
torch.cuda.synchronize()
print("ms per pass:", 200 * (time.perf_counter() - t))
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
  kimimaro_b200.skeletonize(vol.shape, **kw)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(32)
