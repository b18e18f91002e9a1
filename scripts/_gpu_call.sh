mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu --no-strict > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err
python - <<'PY'
import json
r=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print(r["n_gpus"], round(r["ms_per_step"],2), round(r["e2e"]["ms_per_step"],2), r["parity"]["tier_a_identical_to_oracle_in_same_mode"], r["phases_ms"]["pdrf"], r["per_step_ms"])
PY
