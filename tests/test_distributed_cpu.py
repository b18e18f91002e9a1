"""N > 1 host logic on CPU: world_size 2, gloo backend, 127.0.0.1 rendezvous.  Each rank owns a
share of the labels (LPT), rank 0 receives every skeleton through ONE gather of packed buffers."""
import os
import socket

import numpy as np
import pytest


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  import torch
  import torch.distributed as dist
  from kimimaro_b200 import Skeleton, distributed as kd
  dist.init_process_group("gloo", rank=rank, world_size=world)
  counts = {i: 10 * i for i in range(1, 9)}
  mine = kd.make_label_subset(rank, world)(list(counts), counts)
  skels = {}
  for s in mine:
    pts = np.stack([np.arange(s + 2), np.full(s + 2, s), np.zeros(s + 2)], axis=1)
    sk = Skeleton.from_path(pts)
    sk.radii = np.full(s + 2, float(s), np.float32)
    skels[s] = sk
  if rank == 1:      # a label whose components live on two ranks must be merged on rank 0
    skels[100] = Skeleton.from_path(np.array([[0, 0, 9], [1, 0, 9]]))
  else:
    skels[100] = Skeleton.from_path(np.array([[1, 0, 9], [2, 0, 9]]))
  out = kd.gather_skeletons(skels, torch.device("cpu"))
  if rank == 0:
    q.put({k: (v.vertices.copy(), v.edges.copy(), v.radii.copy()) for k, v in out.items()})
  else:
    assert out is None
  dist.barrier()
  dist.destroy_process_group()


def test_gather_world2_gloo():
  pytest.importorskip("torch")
  import torch.multiprocessing as mp
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  out = q.get(timeout=120)
  for p in procs:
    p.join(timeout=120)
    assert p.exitcode == 0
  assert sorted(out) == list(range(1, 9)) + [100]
  for s in range(1, 9):
    v, e, r = out[s]
    assert v.shape == (s + 2, 3) and e.shape == (s + 1, 2) and np.all(r == s)
  v, e, _ = out[100]
  assert v.shape == (3, 3) and e.shape == (2, 2)     # merged + consolidated across ranks


def test_pack_unpack_keeps_negative_and_large_ids():
  """ids are shipped as 64-bit patterns: uint64 ids >= 2^63 and negative ids of signed label dtypes both survive."""
  from kimimaro_b200 import Skeleton, distributed as kd

  def mk():
    return Skeleton.from_path(np.array([[0, 0, 0], [1, 0, 0]]))
  big = {2 ** 63 + 5: mk(), 7: mk()}
  t, v, e, r, tf = kd.pack(big)
  assert sorted(kd.unpack(t, v, e, r, tf, unsigned_ids=True)) == sorted(big)
  neg = {-7: mk(), 3: mk()}
  t, v, e, r, tf = kd.pack(neg)
  assert sorted(kd.unpack(t, v, e, r, tf, unsigned_ids=False)) == sorted(neg)


def _upload_worker(rank, world, port, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  import torch
  import torch.distributed as dist
  from kimimaro_b200 import distributed as kd
  dist.init_process_group("gloo", rank=rank, world_size=world)
  rng = np.random.default_rng(5)
  vol = np.asfortranarray(rng.integers(0, 2 ** 32 - 1, size=(7, 5, 3), dtype=np.uint32))    # 105 voxels: not a multiple of 2
  whole, shape, dtype = kd.upload_sharded(vol, torch.device("cpu"))
  ok = shape == (7, 5, 3) and dtype == np.uint32 and np.array_equal(
    whole.numpy().view(np.uint32), vol.reshape(-1, order="F"))
  # a C-ordered volume is split as it lies and transposed after the all-gather: same Fortran-ordered result
  whole_c, shape_c, _ = kd.upload_sharded(np.ascontiguousarray(vol), torch.device("cpu"))
  ok = ok and shape_c == (7, 5, 3) and np.array_equal(whole_c.numpy().view(np.uint32), vol.reshape(-1, order="F"))
  q.put((rank, bool(ok)))
  dist.barrier()
  dist.destroy_process_group()


def test_upload_sharded_world2_gloo():
  """every rank copies its piece of the volume, the all-gather rebuilds the whole on every rank (ragged last piece)"""
  pytest.importorskip("torch")
  import torch.multiprocessing as mp
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_upload_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  got = dict(q.get(timeout=120) for _ in range(2))
  for p in procs:
    p.join(timeout=120)
    assert p.exitcode == 0
  assert got == {0: True, 1: True}
