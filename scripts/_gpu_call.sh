mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err
python - <<'PY'
import json
r=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print(r["n_gpus"], round(r["ms_per_step"],2), round(r["e2e"]["ms_per_step"],2), r["parity"]["tier_a_identical_to_oracle_in_same_mode"], r["e2e_phases_ms"]["h2d"], r["phases_ms"]["paths"], r["per_step_ms"])
PY
B2T_X=pageable timeout 200 python scripts/phase_times.py 512 2 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pageable', r['pass_ms'], r['phases_ms']['h2d'])"
