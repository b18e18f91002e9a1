#!/usr/bin/env python
"""
bench.py -- voxels/sec of skeletonize() on the 512^3 connectomics-shaped volume (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size 512]

One "step" = one full pass of the hot path (CCL, EDT, per-label TEASAR trace of every connected
component above the dust threshold, skeleton assembly on rank 0) over the whole volume.
  value   voxels/s with the label volume already resident in HBM when the timed region starts
  e2e     the same metric through the public API kimimaro_b200.skeletonize(host ndarray): pinned host
          buffer -> device copy and the device -> host read of every skeleton buffer inside the timed region
The volume is SYNTHETIC (seeded generator, kimimaro_b200/datasets.py): the reference's
benchmarks/connectomics.npy.ckl.gz is crackle-compressed and no decoder exists in this image.

--impl reference times the CPU restatement of the reference (oracle/, all host cores via a fork pool
over labels like kimimaro's parallel mode) on a bounded z-slab of the same volume.
Under torchrun (N > 1) connected components are sharded over the ranks (strong scaling: the volume is
fixed), one NCCL gather of skeleton buffers to rank 0 at the end, time = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ANISOTROPY = (16.0, 16.0, 40.0)        # BASELINE.json configs[1..3] (benchmarks/benchmark.py:27)
ANISOTROPY_1024 = (4.0, 4.0, 40.0)     # BASELINE.json configs[4]
SEED = 0xB2002124
SLAB = (160, 208)          # z-slab of the volume used as the bounded CPU sample (it cuts the lower cap of the soma)
# The CPU legs run the oracle's LITERAL restatement of the reference: binary-heap Dijkstra fields and the heap-ordered
# invalidation of ext/skeletontricks (oracle mode "heap": equal to the reference's compiled extension voxel for voxel,
# tests/test_oracle_cpu.py), not the round-synchronous claim order the oracle offers as the engine's twin -- that one
# is a different, 3-4x faster CPU algorithm and would flatter the CPU side.
CPU_MODE = "heap"


def make_volume(n):
  from kimimaro_b200.datasets import synthetic_tubes
  cache = f"/tmp/b2t_synth_{n}_{SEED:x}.npy"
  if os.path.exists(cache):
    return np.load(cache)
  # git-ignored copy of the generator's output (written by __graft_entry__.build(); it travels with the
  # gpurun snapshot and saves a minute of generation per process on the GPU box)
  packed = os.path.join(ROOT, "oracle", "_cache", f"synth_{n}_{SEED:x}.npz")
  if os.path.exists(packed):
    return np.asfortranarray(np.load(packed)["v"].astype(np.uint32))
  if n == 1024:
    # BASELINE.json configs[4]: 1024^3 = 2x2x2 tiling of the 512^3 volume, label ids offset per tile
    from kimimaro_b200.datasets import tiled
    return tiled(make_volume(512), (2, 2, 2))
  if n >= 512:
    vol = synthetic_tubes((n, n, n), 2124 * (n // 512) ** 3, seed=SEED, anisotropy=ANISOTROPY, soma=True, glia=True)
  else:
    vol = synthetic_tubes((n, n, n), max(8, 2124 * n ** 3 // 512 ** 3), seed=SEED, anisotropy=ANISOTROPY)
  tmp = cache + f".{os.getpid()}.tmp.npy"
  np.save(tmp, vol)
  os.replace(tmp, cache)
  return vol


def anisotropy_of(size):
  return ANISOTROPY_1024 if size == 1024 else ANISOTROPY


def workload_name(vol, size):
  shape = vol.shape
  extra = "one soma + one glia tree, " if size >= 512 else ""
  an = anisotropy_of(size)
  cfg = "configs[4]" if size == 1024 else "configs[2]"
  return (f"synthetic-{size}: {shape[0]}x{shape[1]}x{shape[2]} uint32, {int(len(np.unique(vol)) - 1)} labels, {extra}"
          f"anisotropy {an[0]:g}x{an[1]:g}x{an[2]:g}, DEFAULT_TEASAR_PARAMS, fix_borders (BASELINE.json {cfg} shape; the "
          "reference's connectomics.npy.ckl.gz cannot be decoded in this image)")


def skeleton_digests(sk):
  """per skeleton: first 16 hex digits of sha256(vertices bytes + edges bytes) -- the format of tests/golden/*digest*.json"""
  import hashlib
  out = {}
  for k, s in sk.items():
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(s.vertices).tobytes())
    h.update(np.ascontiguousarray(s.edges).tobytes())
    out[str(k)] = h.hexdigest()[:16]
  return out


def golden_digest(name):
  p = os.path.join(ROOT, "tests", "golden", name)
  if not os.path.exists(p):
    return None
  with open(p) as f:
    return json.load(f)["sha256_16_of_vertices_then_edges"]


class ClockSampler:
  """nvidia-smi clocks / throttle reasons sampled during the timed region."""
  def __init__(self, index=0):
    self.rows, self.proc, self.index = [], None, index

  def start(self):
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
      self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                    "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([c.strip() for c in line.split(",")])

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except Exception:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for r in self.rows:
      try:
        sm.append(float(r[0])); mx.append(float(r[1]))
        for name, val in zip(names, r[3:7]):
          if val.lower().startswith("active"):
            reasons.add(name)
      except Exception:
        pass
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    with open(p) as f:
      return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
  return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def run_reference(args):
  """CPU arm: the oracle port of the reference (literal restatement, heap-ordered invalidation), all host cores.
  value = the WHOLE volume, one pass (the same configuration as the GPU arm: about a minute on 16 cores); the
  --steps / --warmup repetitions run on a bounded z-slab of it and are reported next to it (cpu_baseline.slab)."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  import oracle
  from oracle import teasar
  oracle.build()
  vol = make_volume(args.size)
  an = anisotropy_of(args.size)
  cores = os.cpu_count() or 1
  z0, z1 = (SLAB if args.size >= 256 else (0, args.size))
  sample = np.asfortranarray(vol[:512, :512, z0:z1])
  times = []
  for i in range(args.warmup + args.steps):
    t = time.perf_counter()
    teasar.skeletonize(sample, anisotropy=an, parallel=cores, invalidation_mode=CPU_MODE)
    dt = time.perf_counter() - t
    if i >= args.warmup:
      times.append(dt)
  slab_ms = 1e3 * float(np.mean(times))
  slab_v = sample.size / (slab_ms / 1e3)
  whole = args.size <= 512 and not args.slab_only
  if whole:
    t = time.perf_counter()
    sk = teasar.skeletonize(vol, anisotropy=an, parallel=cores, invalidation_mode=CPU_MODE)
    ms = 1e3 * (time.perf_counter() - t)
    v = vol.size / (ms / 1e3)
    what = f"the whole volume, one pass ({ms / 1e3:.1f} s, {len(sk)} skeletons)"
  else:
    ms, v = slab_ms, slab_v
    what = f"z-slab [{z0}:{z1}) of the volume ({sample.shape[0]}x{sample.shape[1]}x{sample.shape[2]})"
  line = {
    "impl": "reference", "metric": "voxels/sec skeletonized", "value": v, "unit": "voxels/s", "n_gpus": args.gpus,
    "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
    "vs_baseline": None, "dtype": "f32", "data": "synthetic",
    "config": {"workload": workload_name(vol, args.size), "sample": what},
    "cpu_baseline": {"value": v, "unit": "voxels/s", "cores": cores, "kind": "port",
                     "sample": what + "; fork pool over labels, literal restatement (heap-ordered invalidation like "
                                      "ext/skeletontricks)",
                     "slab": {"value": slab_v, "ms_per_step": slab_ms, "steps": args.steps, "warmup": args.warmup,
                              "sample": f"z-slab [{z0}:{z1}) of the same volume"}},
    "e2e": {"value": v, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
  }
  print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_b200(args):
  import torch
  import torch.distributed as dist
  from kimimaro_b200 import _lib, distributed as kdist
  from kimimaro_b200.intake import skeletonize

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  _lib.require_device()

  # data: rank 0 builds (or loads) the cached volume, the others load it afterwards
  if rank == 0:
    vol = make_volume(args.size)
  if world > 1:
    dist.barrier()
  if rank != 0:
    vol = make_volume(args.size)
  shape = vol.shape
  V = vol.size
  AN = anisotropy_of(args.size)
  flat = vol.reshape(-1, order="F")
  pinned = torch.from_numpy(flat.view(np.uint32).view(np.int32)).pin_memory()
  host_view = pinned.numpy().view(np.uint32).reshape(shape, order="F")       # Fortran view of the pinned buffer
  d_labels = pinned.to(dev)
  subset = kdist.make_label_subset(rank, world) if world > 1 else None

  def step_resident(edt_events=None):
    if world > 1:
      return kdist.skeletonize_sharded(shape, device_labels=d_labels, anisotropy=AN, progress=False, edt_events=edt_events)
    return skeletonize(shape, device_labels=d_labels, anisotropy=AN, progress=False, edt_events=edt_events)

  def step_e2e():
    if world > 1:      # every rank copies 1/world of the pinned volume, NCCL all-gather, sharded trace, one gather to rank 0
      return kdist.skeletonize_sharded(host_view, anisotropy=AN, progress=False, in_place=True)
    return skeletonize(host_view, anisotropy=AN, progress=False, in_place=True)

  def sync():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
      torch.cuda.synchronize()

  step_log = {}

  def timed(fn, steps, **kw):
    sync()
    t = time.perf_counter()
    out = None
    marks = []
    for _ in range(steps):
      out = fn(**kw)
      marks.append(time.perf_counter())        # every step ends with a device->host read, so this is its end
    sync()
    dt = time.perf_counter() - t
    step_log[fn.__name__] = [round(1e3 * (b - a), 2) for a, b in zip([t] + marks[:-1], marks)]
    if world > 1:
      tt = torch.tensor([dt], dtype=torch.float64, device=dev)
      dist.all_reduce(tt, op=dist.ReduceOp.MAX)
      dt = float(tt.item())
    return dt, out

  for _ in range(args.warmup):
    step_resident()
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  lib = _lib.lib()
  lib.b2t_launch_count.restype = __import__("ctypes").c_ulonglong
  lib.b2t_launch_count(1)
  edt_events = []
  dt, sk = timed(step_resident, args.steps, edt_events=edt_events)
  launches = int(lib.b2t_launch_count(0)) // max(args.steps, 1)
  clocks = sampler.stop() if rank == 0 else None
  ms = 1e3 * dt / args.steps
  value = V / (ms / 1e3)

  # K1 duration from CUDA events recorded around b2t_edt on its launch stream inside the timed steps
  torch.cuda.synchronize()
  edt_ms = float(np.mean([a.elapsed_time(b) for a, b in edt_events])) if edt_events else None

  # end to end through the public API with host buffers
  for _ in range(min(args.warmup, 1)):
    step_e2e()
  dte, sk_e = timed(step_e2e, args.steps)
  mse = 1e3 * dte / args.steps
  d2h = 0
  if rank == 0 and sk_e:
    d2h = int(sum(s.vertices.nbytes + s.edges.nbytes + s.radii.nbytes for s in sk_e.values()))

  # per-phase breakdowns (extra steps with synchronising laps, not part of the metric)
  tme = {}
  bshard = (rank, world, None) if world > 1 else None          # the six faces of the border targets split over the ranks, as in the timed steps
  skeletonize(host_view, anisotropy=AN, progress=False, in_place=True, label_subset=subset, timings=tme, border_shard=bshard)
  e2e_phases = {k: round(1e3 * v, 3) for k, v in tme.items() if isinstance(v, float)}
  tm = {}
  skeletonize(shape, device_labels=d_labels, anisotropy=AN, progress=False, label_subset=subset, timings=tm, border_shard=bshard)
  phases = {k: round(1e3 * v, 3) for k, v in tm.items() if isinstance(v, float)}
  phases_per_rank = None
  if world > 1:              # what every rank spends where (its share of the labels): the limiter of the strong scaling
    phases_per_rank = [None] * world
    dist.all_gather_object(phases_per_rank, phases)

  # parity with the REFERENCE's invalidation order, measured on this very run (outside the timed region): the timed
  # steps' skeletons against the digest of the oracle's literal heap order (== the reference's compiled extension), then
  # one pass in the engine's strict mode, which must hit all of them
  parity = None
  mode, window = _lib.invalidation_mode()
  heap_gold = golden_digest("synth512_oracle_digest_heap.json") if args.size == 512 else None
  if heap_gold is not None:
    same_gold = golden_digest("synth512_oracle_digest_window1.json")
    strict = None
    if not args.no_strict:
      _lib.set_invalidation_mode("strict")
      try:
        sync()
        t = time.perf_counter()
        sk_s = step_resident()
        sync()
        strict_ms = 1e3 * (time.perf_counter() - t)
      finally:
        _lib.set_invalidation_mode(mode, window)
      if rank == 0:
        ds = skeleton_digests(sk_s)
        strict = {"identical_to_reference_order": int(sum(ds.get(k) == h for k, h in heap_gold.items())),
                  "of": len(heap_gold), "ms_per_step": strict_ms}
    if rank == 0:
      dg = skeleton_digests(sk)
      parity = {"mode": mode + (f":{window:g}" if mode == "window" else ""),
                "identical_to_reference_order": int(sum(dg.get(k) == h for k, h in heap_gold.items())),
                "of": len(heap_gold),
                "tier_a_identical_to_oracle_in_same_mode": (int(sum(dg.get(k) == h for k, h in same_gold.items()))
                                                            if (same_gold and mode == "window" and window == 1.0) else None),
                "strict": strict,
                "reference_order": "oracle mode 'heap' == ext/skeletontricks compiled (tests/test_oracle_cpu.py), "
                                   "tests/golden/synth512_oracle_digest_heap.json"}

  if rank == 0:
    peak, how = peaks()
    label_bytes = 4
    alg = (3 * label_bytes + 20) * V                     # (3L+20) B/voxel, three separable passes
    roof = None
    if edt_ms:
      ach = alg / (edt_ms * 1e-3) / 1e9
      traffic = None
      tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")      # dram bytes of one K1 call from an ncu capture
      if os.path.exists(tpath) and args.size == 512:
        with open(tpath) as tf:
          traffic = float(json.load(tf)["dram_bytes_read_plus_write"])
      roof = {"kernel": "b2t_edt_ws (K1: edt_pass_x_v2 + 2x [edt_pass_col_stencil + edt_pass_col_fh3_range])", "bound": "hbm", "achieved": ach, "peak": peak,
              "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "peak_source": how,
              "algorithmic_bytes": alg, "ms": edt_ms}
    # bounded CPU sample: the oracle port, one core, on the slab
    cpu = None
    if not args.no_cpu:
      import oracle
      from oracle import teasar
      oracle.build()
      z0, z1 = (SLAB if args.size >= 256 else (0, args.size))
      sample = np.asfortranarray(vol[:512, :512, z0:z1])
      t = time.perf_counter()
      teasar.skeletonize(sample, anisotropy=AN, invalidation_mode=CPU_MODE)
      cdt = time.perf_counter() - t
      cpu = {"value": sample.size / cdt, "unit": "voxels/s", "cores": 1, "kind": "port",
             "sample": f"z-slab [{z0}:{z1}) of the same volume ({sample.shape[0]}x{sample.shape[1]}x{sample.shape[2]}), {cdt:.1f} s, "
                       f"literal restatement (heap-ordered invalidation like ext/skeletontricks)"}
    line = {
      "metric": "voxels/sec skeletonized", "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
      "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
      "dtype": "f32", "data": "synthetic",
      "config": {"workload": workload_name(vol, args.size),
                 "l2": "inputs and work fields (>3 GB) exceed the 126 MB L2", "skeletons": len(sk) if sk else 0,
                 "parallelism": f"labels sharded over {world} rank(s)"},
      "e2e": {"value": V / (mse / 1e3), "unit": "voxels/s", "h2d_bytes_per_step": int(flat.nbytes),
              "d2h_bytes_per_step": d2h, "ms_per_step": mse},
      "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "parity": parity,
      "phases_ms": phases, "phases_ms_per_rank": phases_per_rank,
      "e2e_phases_ms": e2e_phases, "per_step_ms": step_log,
    }
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=3)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--size", type=int, default=512)
  ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
  ap.add_argument("--slab-only", action="store_true", help="--impl reference: skip the whole-volume pass")
  ap.add_argument("--no-strict", action="store_true", help="skip the strict-mode pass (parity.strict)")
  args = ap.parse_args()
  if args.impl == "reference":
    run_reference(args)
  else:
    run_b200(args)


if __name__ == "__main__":
  main()
