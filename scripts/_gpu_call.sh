set -x
mkdir -p gpurun_out
: > gpurun_out/phase_ab.jsonl
for rep in 1 2; do
  for v in t256_minb2 t256_minb3 minb2 minb1 t384_minb1; do
    L=$PWD/kimimaro_b200/_variants/$v.so
    B2T_LIB=$L B2T_X=$v timeout 300 python scripts/phase_times.py 512 3 >> gpurun_out/phase_ab.jsonl 2>> gpurun_out/phase_ab.err
  done
done

python - <<'PY'
import json
for l in open("gpurun_out/phase_ab.jsonl"):
  r = json.loads(l)
  print(r["env"].get("B2T_X"), r["pass_ms"], r["phases_ms"]["paths"], r.get("identical_to_oracle_same_mode"), r.get("label_us"), [ (s["job"], s["us"]) for s in r.get("slowest_labels", [])[:3]])
PY
