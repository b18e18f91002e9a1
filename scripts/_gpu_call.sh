mkdir -p gpurun_out
: > gpurun_out/phase_ab.jsonl
for v in 1 0; do
  B2T_STAGED_UPLOAD=$v B2T_X=staged$v timeout 300 python scripts/phase_times.py 512 3 >> gpurun_out/phase_ab.jsonl 2>> gpurun_out/phase_ab.err
done
tail -3 gpurun_out/phase_ab.err
python - <<'PY'
import json
for l in open("gpurun_out/phase_ab.jsonl"):
  r = json.loads(l); ph = r["phases_ms"]
  print(r["env"].get("B2T_X"), r["pass_ms"], ph.get("h2d"), ph.get("ccl"), r.get("identical_to_oracle_same_mode"))
PY
python - <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0, ".")
import kimimaro_b200
from bench import make_volume
vol = make_volume(512)
c = np.ascontiguousarray(vol)
kimimaro_b200.skeletonize(c, anisotropy=(16,16,40), progress=False)
t=time.perf_counter(); a = kimimaro_b200.skeletonize(c, anisotropy=(16,16,40), progress=False); torch.cuda.synchronize(); print("C-order pass ms", 1e3*(time.perf_counter()-t), len(a))
t=time.perf_counter(); b = kimimaro_b200.skeletonize(vol, anisotropy=(16,16,40), progress=False); torch.cuda.synchronize(); print("F-order pass ms", 1e3*(time.perf_counter()-t), len(b))
print("equal", sorted(a)==sorted(b) and all(np.array_equal(a[k].vertices,b[k].vertices) and np.array_equal(a[k].edges,b[k].edges) for k in a))
PY
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -1
