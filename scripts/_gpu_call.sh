mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
python - <<'PY'
import json
r=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print(r["n_gpus"], r["steps"], round(r["ms_per_step"],2), round(r["e2e"]["ms_per_step"],2), r["parity"]["identical_to_reference_order"], r["parity"]["tier_a_identical_to_oracle_in_same_mode"], r["parity"]["strict"]["identical_to_reference_order"], round(r["parity"]["strict"]["ms_per_step"]), r["roofline"]["frac"], r["phases_ms"]["paths"], r["cpu_baseline"]["value"], r["gpu_launches"], r["clocks"])
PY
