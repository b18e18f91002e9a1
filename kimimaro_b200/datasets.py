"""
Seeded synthetic label volumes shaped like BASELINE.json's configs.

The reference's benchmark volume (benchmarks/connectomics.npy.ckl.gz) is crackle-compressed and
no decoder exists in this image, so configs 2-5 are generated: branching random-walk "neurites"
painted into a zero volume in seed order (later tubes never overwrite earlier ones), optionally one
soma-sized ellipsoid (crosses both soma thresholds of intake.py:52-53 at 16x16x40 nm) and one large
branching "glia" tree.  Every number measured on these volumes is labelled "synthetic".
"""
import numpy as np


def _ball_offsets(r, zscale):
  rz = max(1, int(np.ceil(r * zscale)))
  ri = int(np.ceil(r))
  xs, ys, zs = np.meshgrid(np.arange(-ri, ri + 1), np.arange(-ri, ri + 1), np.arange(-rz, rz + 1), indexing="ij")
  m = xs * xs + ys * ys + (zs / zscale) ** 2 <= r * r
  return np.stack([xs[m], ys[m], zs[m]], axis=1).astype(np.int64)


def _walk(rng, start, direction, steps, shape, zscale, momentum=0.9):
  """26-step-ish random walk with momentum; returns integer centres inside the volume."""
  noise = rng.normal(size=(steps, 3))
  pts = np.empty((steps, 3), dtype=np.float64)
  p = np.array(start, dtype=np.float64)
  d = np.array(direction, dtype=np.float64)
  d /= np.linalg.norm(d) + 1e-12
  scale = np.array([1.0, 1.0, zscale])
  n = 0
  hi = np.array(shape, dtype=np.float64) - 1
  for i in range(steps):
    d = momentum * d + (1.0 - momentum) * 3.0 * noise[i]
    d /= np.linalg.norm(d) + 1e-12
    p = p + d * scale
    if (p < 0).any() or (p > hi).any():
      break
    pts[n] = p
    n += 1
  return pts[:n], d


def _paint(vol_flat, shape, centres, offs, label):
  if centres.shape[0] == 0:
    return
  c = np.round(centres[::2]).astype(np.int64)
  v = (c[:, None, :] + offs[None, :, :]).reshape(-1, 3)
  ok = (v >= 0).all(axis=1) & (v[:, 0] < shape[0]) & (v[:, 1] < shape[1]) & (v[:, 2] < shape[2])
  v = v[ok]
  idx = v[:, 0] + shape[0] * (v[:, 1] + shape[1] * v[:, 2])
  idx = np.unique(idx)
  idx = idx[vol_flat[idx] == 0]
  vol_flat[idx] = label


def synthetic_tubes(shape, n_labels, seed, anisotropy=(16.0, 16.0, 40.0), soma=False, glia=False,
                    dtype=np.uint32, branch_prob=0.006):
  """Returns a Fortran-ordered label volume with labels 1..n_labels (some may end up empty or split)."""
  shape = tuple(int(s) for s in shape)
  rng = np.random.default_rng(seed)
  vol = np.zeros(shape, dtype=dtype, order="F")
  flat = vol.reshape(-1, order="F")
  zscale = float(anisotropy[0]) / float(anisotropy[2])
  label = 0
  if soma:
    # physical radius 3600 nm -> 225 x 225 x 90 voxel semi-axes at 16,16,40
    label += 1
    c = np.array([shape[0] * 0.5, shape[1] * 0.5, shape[2] * 0.5])
    r = 3600.0
    xs, ys, zs = np.ogrid[:shape[0], :shape[1], :shape[2]]
    m = ((xs - c[0]) * anisotropy[0]) ** 2 + ((ys - c[1]) * anisotropy[1]) ** 2 + ((zs - c[2]) * anisotropy[2]) ** 2 <= r * r
    vol[m] = label
    # a few thick dendrites leaving the soma
    for _ in range(5):
      d = rng.normal(size=3)
      pts, _ = _walk(rng, c, d, 700, shape, zscale, momentum=0.97)
      _paint(flat, shape, pts, _ball_offsets(9.0, zscale), label)
  if glia:
    label += 1
    todo = [(rng.uniform(0.3, 0.7, size=3) * np.array(shape), rng.normal(size=3), 400, 6.0)]
    painted = 0
    while todo and painted < 60:
      start, d, steps, r = todo.pop()
      pts, dend = _walk(rng, start, d, steps, shape, zscale, momentum=0.8)
      _paint(flat, shape, pts, _ball_offsets(r, zscale), label)
      painted += 1
      if pts.shape[0] > 10 and r > 2.5:
        for _ in range(3):
          j = rng.integers(5, pts.shape[0])
          todo.append((pts[j], rng.normal(size=3), int(steps * 0.7), r * 0.8))
  while label < n_labels:
    label += 1
    r = rng.uniform(3.0, 9.0)
    steps = int(rng.integers(150, 600))
    start = rng.uniform(0, 1, size=3) * (np.array(shape) - 1)
    pts, _ = _walk(rng, start, rng.normal(size=3), steps, shape, zscale)
    offs = _ball_offsets(r, zscale)
    _paint(flat, shape, pts, offs, label)
    # side branches
    nb = rng.binomial(max(pts.shape[0], 1), branch_prob)
    for _ in range(int(nb)):
      if pts.shape[0] < 5:
        break
      j = int(rng.integers(0, pts.shape[0]))
      rb = max(2.0, r * rng.uniform(0.5, 0.9))
      bpts, _ = _walk(rng, pts[j], rng.normal(size=3), int(rng.integers(40, 250)), shape, zscale)
      _paint(flat, shape, bpts, _ball_offsets(rb, zscale), label)
  return vol


def sphere(n=64, radius=24, dtype=np.uint8):
  """BASELINE.json config 0: 64^3 single-label sphere."""
  c = n // 2
  xs, ys, zs = np.ogrid[:n, :n, :n]
  return np.asfortranarray(((xs - c) ** 2 + (ys - c) ** 2 + (zs - c) ** 2 <= radius * radius).astype(dtype))


def tiled(vol, reps=(2, 2, 2)):
  """BASELINE.json config 4: tile a volume, offsetting non-zero labels by t*2^20 per tile."""
  sx, sy, sz = vol.shape
  out = np.zeros((sx * reps[0], sy * reps[1], sz * reps[2]), dtype=np.uint32, order="F")
  t = 0
  for k in range(reps[2]):
    for j in range(reps[1]):
      for i in range(reps[0]):
        blk = vol.astype(np.uint32)
        blk = np.where(blk != 0, blk + np.uint32(t << 20), 0)
        out[i * sx:(i + 1) * sx, j * sy:(j + 1) * sy, k * sz:(k + 1) * sz] = blk
        t += 1
  return out
