"""Development: the GEMV kernel families head to head on shapes at the edge of the streaming kernel's domain (few row
tiles: k/v projections, tensor-parallel shards).  Decides which of the older families still earn their place."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_bnb_fp4_ext as ext  # noqa: E402
from torch_bnb_fp4_b200 import _lib  # noqa: E402
from tools.microbench import graph_time  # noqa: E402

dev = torch.device("cuda:0")
code = torch.tensor(ext.BNB_FP4_CODE, device=dev)
# (round 2: the stream-K families gemv_i8 / gemv_tma / gemv_imma lost on every shape below - see
# profiles/r02_gemv_family_sweep.log - and were removed; what is left to compare is the generic kernel)
FAM = {"stream": 0, "generic": _lib.FLAG_FORCE_GENERIC}
for shp in sys.argv[1:] or ["128x4096", "256x4096", "512x4096", "768x4096", "1024x4096", "512x14336", "2048x768", "64x2048",
                            "1024x8192", "128x8192", "3584x8192"]:
    N, K = map(int, shp.split("x"))
    nrot = 32
    Ws = [torch.randint(0, 256, (N * K // 2, 1), dtype=torch.uint8, device=dev) for _ in range(nrot)]
    ams = [torch.rand(N * K // 64, device=dev) * 0.1 + 0.01 for _ in range(nrot)]
    for b in (1, 8):
        x = torch.randn(b, K, device=dev).bfloat16()
        row = []
        for name, fl in FAM.items():
            fn = lambda i: ext.gemv_fp4_bias(x, Ws[i], ams[i], code, 64, ext.bfloat16, [N, K], None, None, fl)  # noqa: E731
            row.append(f"{name} {graph_time(fn, nrot):7.2f}")
        print(f"{shp:>11s} b={b}: " + " | ".join(row) + " us", flush=True)
