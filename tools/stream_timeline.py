"""Debug: per-warp phase timeline (globaltimer ns) of a CHAIN of gemv_stream launches (each layer's input
is the previous layer's output) replayed from a CUDA graph.  Needs a build with
FP4_B200_NVCC_EXTRA=-DFP4_STREAM_TIMELINE.  Usage: stream_timeline.py N K [nlaunch]   (N == K chains; else
the same x feeds every launch)"""
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
ams = [torch.rand(N * K // 64, device=dev) * 0.02 + 0.01 for _ in range(NL)]
x0 = torch.randn(1, K, device=dev).bfloat16()
stride = 148 * KW * 8
buf = torch.zeros(stride * (NL * 4 + 8), dtype=torch.int64, device=dev)


def chain():
    x = x0
    for i in range(NL):
        y = ext.gemv_fp4(x, Ws[i], ams[i], code, 64, ext.bfloat16, [N, K])
        if N == K:
            x = y
    return y


s = torch.cuda.Stream()
with torch.cuda.stream(s):
    chain()
    torch.cuda.synchronize()
    lib.fp4_b200_debug_stream_timeline.argtypes = [ctypes.c_void_p]
    lib.fp4_b200_debug_stream_timeline(buf.data_ptr())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        chain()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
raw = buf.cpu().numpy().reshape(-1, 148 * KW, 8)[:NL]
t00 = raw[raw > 0].min()
names = ["start", "ring issued", "dep-wait done", "x staged", "first data", "loop done", "all warps done", "end"]
for li in range(NL):
    r = raw[li]
    print(f"launch {li}: active warps {(r[:, 0] > 0).sum()}")
    for j, nm in enumerate(names):
        col = r[:, j][r[:, j] > 0]
        if col.size:
            print(f"   {nm:14s} min {col.min() - t00:8d}  mean {col.mean() - t00:10.0f}  max {col.max() - t00:8d} ns  (n={col.size})")
