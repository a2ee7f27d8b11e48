"""Development microbenchmark: per-kernel device time with rotating weight buffers (> 2x L2) so the
numbers are HBM, not L2.  Not the judged bench (that is bench.py)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_bnb_fp4_ext as ext  # noqa: E402
from torch_bnb_fp4_b200 import _lib  # noqa: E402

PEAK = 6557.1
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass


def timeit(fn, nrot, reps=5):
    for i in range(min(nrot, 8)):
        fn(i)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(nrot):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / nrot * 1e3)
    return best  # us per call


def graph_time(fn, nrot, reps=5):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(min(nrot, 4)):
            fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(nrot):
                fn(i)
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / nrot * 1e3)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--batch", type=int, nargs="+", default=[1])
    ap.add_argument("--generic", action="store_true")
    ap.add_argument("--no-dequant", action="store_true")
    ap.add_argument("--shapes", nargs="*", default=["4096x4096", "1024x4096", "14336x4096", "4096x14336",
                                                     "8192x8192", "28672x8192"])
    a = ap.parse_args()
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[a.dtype]
    st = {"bf16": ext.bfloat16, "fp16": ext.float16, "fp32": ext.float32}[a.dtype]
    dev = torch.device("cuda:0")
    code = torch.tensor(ext.BNB_FP4_CODE, device=dev)
    flags = _lib.FLAG_FORCE_GENERIC if a.generic else 0
    for shp in a.shapes:
        N, K = map(int, shp.split("x"))
        wbytes = N * K // 2
        nrot = max(4, int(2.5 * 128e6 / (wbytes * 1.125)) + 1)
        nrot = min(nrot, 128)
        Ws = [torch.randint(0, 256, (wbytes, 1), dtype=torch.uint8, device=dev) for _ in range(nrot)]
        ams = [torch.rand(N * K // 64, device=dev) * 0.1 + 0.01 for _ in range(nrot)]
        for b in a.batch:
            x = torch.randn(b, K, device=dev).to(dt)
            fn = lambda i: ext.gemv_fp4_bias(x, Ws[i], ams[i], code, 64, st, [N, K], None, None, flags)  # noqa: E731
            us = graph_time(fn, nrot)
            byts = wbytes + N * K // 64 * 4 + b * K * x.element_size() + b * N * x.element_size()
            print(f"gemv {shp} b={b} {a.dtype}: {us:8.2f} us  {byts / us / 1e3:8.1f} GB/s  "
                  f"{byts / us / 1e3 / PEAK * 100:5.1f}% of measured {PEAK:.0f}", flush=True)
        if not a.no_dequant:
            fn = lambda i: ext.dequantize_fp4(Ws[i], ams[i], 64, N, K, st)  # noqa: E731
            us = graph_time(fn, min(nrot, 16))
            byts = wbytes + N * K // 64 * 4 + N * K * x.element_size()
            print(f"dequant {shp} {a.dtype}: {us:8.2f} us  {byts / us / 1e3:8.1f} GB/s  "
                  f"{byts / us / 1e3 / PEAK * 100:5.1f}% of measured", flush=True)
        del Ws, ams
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
