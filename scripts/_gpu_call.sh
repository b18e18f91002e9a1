mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-strict > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 300 gpurun_out/bench_n2.err
python - <<'PY'
import json
r=json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print(r["n_gpus"], round(r["ms_per_step"],2), round(r["e2e"]["ms_per_step"],2), r["parity"]["tier_a_identical_to_oracle_in_same_mode"], [round(p["total"],1) for p in r["phases_ms_per_rank"]], r["per_step_ms"])
PY
