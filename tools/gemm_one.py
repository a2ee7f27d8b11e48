"""Development: one shape of the fused GEMM, a few launches (for ncu).  usage: gemm_one.py M [N K]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_bnb_fp4_ext as ext  # noqa: E402

M = int(sys.argv[1])
N, K = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (28672, 8192)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
packed = torch.randint(0, 256, (N * K // 2, 1), dtype=torch.uint8, device=dev, generator=g)
absmax = torch.rand(N * K // 64, device=dev, generator=g) * 0.02 + 0.01
code = torch.tensor(ext.BNB_FP4_CODE, device=dev)
x = torch.randn(M, K, device=dev, generator=g).bfloat16()
for _ in range(4):
    y = ext.gemm_fp4(x, packed, absmax, code, N, K, 64)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    y = ext.gemm_fp4(x, packed, absmax, code, N, K, 64)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 200
print(f"M={M} {N}x{K}: {us:.1f} us  {2.0 * M * N * K / us / 1e6:.0f} TFLOP/s")
