mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_edt_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 90 python - <<'PY'
import ctypes, sys, json
sys.path.insert(0, ".")
import numpy as np, torch
from kimimaro_b200 import ops, _lib
import bench
vol = bench.make_volume(512)
L = _lib.lib(); L.b2t_edt_config_xpass.restype = ctypes.c_longlong; L.b2t_edt_config_xpass.argtypes = [ctypes.c_int]
d = ops.to_device_f(vol)
out = torch.empty(vol.size, dtype=torch.float32, device="cuda")
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
res = {}
for form in (0, 1, 2, 0, 1, 2):
  L.b2t_edt_config_xpass(form)
  for _ in range(3): ops.edt(d, vol.shape, (16, 16, 40), False, out=out)
  ts = []
  for _ in range(9):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.edt(d, vol.shape, (16, 16, 40), False, out=out); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  res.setdefault(form, []).append(round(float(np.median(ts)), 4))
print(json.dumps({"k1_ms_median_by_xpass_form": res, "tma_launches": int(L.b2t_edt_config_xpass(-1))}))
PY
