set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 8 --warmup 3 --no-cpu --no-strict > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; tail -c 300 gpurun_out/bench_n$n.err
done
timeout 300 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -2
python - <<'PY'
import json
for n in (2,4,8):
  f=f"gpurun_out/bench_n{n}.json"
  try:
    r=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, r["n_gpus"], round(r["ms_per_step"],2), round(r["e2e"]["ms_per_step"],2), r["parity"]["tier_a_identical_to_oracle_in_same_mode"]); print([round(p["total"],1) for p in r["phases_ms_per_rank"]]); print(r["per_step_ms"])
  except Exception as e: print(f, "ERR", e)
PY
