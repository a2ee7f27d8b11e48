"""Prefill sweep (BASELINE config 5): dequant-fused tcgen05 GEMM vs new-dequant + cuBLAS on one FP4 weight.
Usage: gemm_sweep.py [N K] [--m 1 16 64 ...]   (default 28672 x 8192, i.e. a Llama-3-70B gate/up projection)"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_bnb_fp4_ext as ext  # noqa: E402

PK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best  # ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("shape", nargs="*", type=int, default=[28672, 8192])
    ap.add_argument("--m", type=int, nargs="+", default=[16, 32, 64, 128, 256, 512, 1024, 2048, 4096])
    ap.add_argument("--dtype", default="bf16")
    a = ap.parse_args()
    N, K = a.shape
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[a.dtype]
    st = {"bf16": ext.bfloat16, "fp16": ext.float16}[a.dtype]
    dev = torch.device("cuda:0")
    packed = torch.randint(0, 256, (N * K // 2, 1), dtype=torch.uint8, device=dev)
    absmax = torch.rand(N * K // 64, device=dev) * 0.02 + 0.005
    for M in a.m:
        x = torch.randn(M, K, device=dev).to(dt)
        fused = lambda: ext.gemm_fp4(x, packed, absmax, None, N, K, 64)  # noqa: E731
        lib = lambda: torch.nn.functional.linear(x, ext.dequantize_fp4(packed, absmax, 64, N, K, st))  # noqa: E731
        y0, y1 = fused(), lib()
        err = (y0.float() - y1.float()).abs().max().item() / y1.float().abs().max().item()
        tf, tl = timeit(fused), timeit(lib)
        fl = 2.0 * M * N * K
        print(f"M={M:5d}  fused {tf * 1e3:9.1f} us {fl / tf / 1e9:8.1f} TFLOP/s ({fl / tf / 1e9 / PK['bf16_tflops'] * 100:5.1f}% of measured "
              f"{PK['bf16_tflops']:.0f})   dequant+cuBLAS {tl * 1e3:9.1f} us {fl / tl / 1e9:8.1f} TFLOP/s   speed-up {tl / tf:5.2f}x   "
              f"max rel diff {err:.1e}", flush=True)


if __name__ == "__main__":
    main()
