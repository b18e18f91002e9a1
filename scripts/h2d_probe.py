import time, numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n = 512**3
pinned = torch.empty(n, dtype=torch.int32).pin_memory()
view = pinned.numpy().view(np.uint32).reshape((512,512,512), order="F")
flat = view.reshape(-1, order="F")
t2 = torch.from_numpy(flat.view(np.uint32))
print("numpy view shares memory:", np.shares_memory(flat, pinned.numpy()), "is_pinned(from_numpy uint32):", t2.is_pinned(), "is_pinned(int32 view):", torch.from_numpy(flat.view(np.int32)).is_pinned())
for name, src in (("pinned tensor", pinned), ("from_numpy uint32 view", t2), ("from_numpy int32 view", torch.from_numpy(flat.view(np.int32))), ("pageable", torch.empty(n, dtype=torch.int32))):
  for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); d = src.cuda(non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
  print(f"{name:28s} {dt*1e3:7.2f} ms  {n*4/dt/1e9:6.1f} GB/s")
