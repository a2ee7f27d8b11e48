"""Debug: per-warp phase timeline (globaltimer ns) of a chain of gemv_i8 launches replayed from a CUDA
graph.  Needs a build with FP4_B200_NVCC_EXTRA=-DFP4_I8_TIMELINE.  Usage: i8_timeline.py N K [nlaunch]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_bnb_fp4_ext as ext  # noqa: E402
from torch_bnb_fp4_b200._lib import lib  # noqa: E402

N, K = (int(v) for v in sys.argv[1:3])
NL = int(sys.argv[3]) if len(sys.argv) > 3 else 8
KW = int(os.environ.get("KW", "16"))
dev = torch.device("cuda:0")
code = torch.tensor(ext.BNB_FP4_CODE, device=dev)
Ws = [torch.randint(0, 256, (N * K // 2, 1), dtype=torch.uint8, device=dev) for _ in range(NL)]
ams = [torch.rand(N * K // 64, device=dev) * 0.1 + 0.01 for _ in range(NL)]
x = torch.randn(1, K, device=dev).bfloat16()
stride = 148 * 4 * KW * 8
buf = torch.zeros(stride * (NL * 3 + 8), dtype=torch.int64, device=dev)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for i in range(2):
        ext.gemv_fp4(x, Ws[i], ams[i], code, 64, ext.bfloat16, [N, K])
    torch.cuda.synchronize()
    lib.fp4_b200_debug_timeline.argtypes = [ctypes.c_void_p]
    lib.fp4_b200_debug_timeline(buf.data_ptr())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for i in range(NL):
            ext.gemv_fp4(x, Ws[i], ams[i], code, 64, ext.bfloat16, [N, K])
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
raw = buf.cpu().numpy().reshape(-1, 148 * 4 * KW, 8)[:NL]
t00 = raw[raw > 0].min()
names = ["start", "issued", "dep-wait done", "x staged", "first data", "loop done", "flush done"]
for li in range(NL):
    r = raw[li]
    act = r[:, 0] > 0
    print(f"launch {li}: active warps {act.sum()}")
    for j, nm in enumerate(names):
        col = r[:, j][r[:, j] > 0]
        if col.size:
            print(f"   {nm:14s} min {col.min() - t00:8d}  mean {col.mean() - t00:10.0f}  max {col.max() - t00:8d} ns  (n={col.size})")
