"""Debug: per-warp phase timeline of the GEMV kernel (needs a build with -DFP4_GEMV_TIMELINE)."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_bnb_fp4_ext as ext
from torch_bnb_fp4_b200 import ext as E

N, K = (int(v) for v in (sys.argv[1:3] if len(sys.argv) > 2 else (4096, 4096)))
dev = torch.device("cuda:0")
code = torch.tensor(ext.BNB_FP4_CODE, device=dev)
W = torch.randint(0, 256, (N * K // 2, 1), dtype=torch.uint8, device=dev)
am = torch.rand(N * K // 64, device=dev) * 0.1 + 0.01
x = torch.randn(1, K, device=dev).bfloat16()
for _ in range(5):
    y = ext.gemv_fp4(x, W, am, code, 64, ext.bfloat16, [N, K])
torch.cuda.synchronize()
ws = list(E._workspaces.values())[0]
units = (N // 16) * (K // 64)
grid = min(148 * 2, (units + 63) // 64)
Wn = grid * 8
off = 256 * 1024 + Wn * 2 * 16 * 8 * 4
raw = ws[off: off + Wn * 64].cpu().numpy().view(np.int64).reshape(Wn, 8)
t0 = raw[:, 0].astype(np.float64)
print("grid", grid, "warps", Wn)
base = raw[:, 0].min()
for nm, i, j in [("issue loads", 0, 1), ("stage x", 1, 2), ("main loop", 2, 3), ("flush", 3, 4), ("total", 0, 4)]:
    d = (raw[:, j] - raw[:, i]).astype(np.float64)
    print(f"{nm:12s} cycles: mean {d.mean():8.0f}  min {d.min():8.0f}  max {d.max():8.0f}")
print("start spread (cycles, same-SM clocks only comparable):", (raw[:, 0] - base).max())
g = raw[:, 6]
print("globaltimer end spread ns:", g.max() - g.min())
