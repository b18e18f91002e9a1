set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -1
timeout 600 python -m pytest tests/test_skeletonize_gpu.py -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
: > gpurun_out/phase_ab.jsonl
for rep in 1 2; do
  for v in default rr_global minb2 minb1 batch2 minb2_batch2; do
    if [ $v = default ]; then L=$PWD/kimimaro_b200/libb2t.so; else L=$PWD/kimimaro_b200/_variants/$v.so; fi
    B2T_LIB=$L B2T_X=$v timeout 300 python scripts/phase_times.py 512 3 >> gpurun_out/phase_ab.jsonl 2>> gpurun_out/phase_ab.err
  done
done
B2T_LIB=$PWD/kimimaro_b200/_variants/prof.so timeout 300 python scripts/trace_prof.py > gpurun_out/trace_prof.jsonl 2> gpurun_out/trace_prof.err
tail -3 gpurun_out/trace_prof.err
python - <<'PY'
import json
for l in open("gpurun_out/phase_ab.jsonl"):
  r = json.loads(l)
  print(r["env"].get("B2T_X"), r["pass_ms"], r["phases_ms"]["paths"], r.get("identical_to_oracle_same_mode"), [ (s["job"], s["us"]) for s in r.get("slowest_labels", [])[:3]])
PY
