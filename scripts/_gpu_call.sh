set -x
mkdir -p gpurun_out
: > gpurun_out/phase_ab.jsonl
for rep in 1 2; do
  for v in default edf_solo_lanes; do
    if [ $v = default ]; then L=$PWD/kimimaro_b200/libb2t.so; else L=$PWD/kimimaro_b200/_variants/$v.so; fi
    B2T_LIB=$L B2T_X=$v timeout 300 python scripts/phase_times.py 512 3 >> gpurun_out/phase_ab.jsonl 2>> gpurun_out/phase_ab.err
  done
done
tail -3 gpurun_out/phase_ab.err
python - <<'PY'
import json
for l in open("gpurun_out/phase_ab.jsonl"):
  r = json.loads(l); ph = r["phases_ms"]
  print(r["env"].get("B2T_X"), r["pass_ms"], {k: ph.get(k) for k in ("find_root","daf","paths","soma","soma_daf")}, r.get("identical_to_oracle_same_mode"))
PY
