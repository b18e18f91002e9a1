set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err
python - <<'PY'
import json
r=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print(r["n_gpus"], round(r["ms_per_step"],2), round(r["e2e"]["ms_per_step"],2), r["parity"]["identical_to_reference_order"], r["parity"]["strict"], r["gpu_launches"], r["roofline"]["frac"], r["roofline"]["ms"], r["cpu_baseline"]["value"]); print(r["phases_ms"]); print(r["per_step_ms"]); print(r["clocks"])
PY
