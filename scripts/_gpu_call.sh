set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/gpu_tests_n2.log 2>&1; tail -3 gpurun_out/gpu_tests_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu --no-strict > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 400 gpurun_out/bench_n2.err
B2T_AN=4,4,40 B2T_LIB=$PWD/kimimaro_b200/_variants/prof.so timeout 600 python scripts/trace_prof.py 512 > gpurun_out/trace_prof_4440.jsonl 2> gpurun_out/trace_prof_4440.err; tail -3 gpurun_out/trace_prof_4440.err; head -c 1500 gpurun_out/trace_prof_4440.jsonl
python - <<'PY'
import json
for f in ("gpurun_out/bench_n2.json",):
  try:
    r=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, r["n_gpus"], round(r["ms_per_step"],2), round(r["e2e"]["ms_per_step"],2), r["parity"]["tier_a_identical_to_oracle_in_same_mode"]); print(r.get("phases_ms_per_rank")); print(r["per_step_ms"])
  except Exception as e: print(f, "ERR", e)
PY
