"""Small deterministic workload for compute-sanitizer (memcheck / racecheck / synccheck, one tool per run): every
kernel of libfp4_b200 once or twice on small shapes - the streaming GEMV (plain, half units, ALIGNED, nested, grouped,
gated and residual epilogues), the generic GEMV, the tcgen05 GEMM (two token-tile sizes), dequant, quantiser.
    compute-sanitizer --tool racecheck python tools/sanitize_case.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_bnb_fp4_ext as ext  # noqa: E402
from torch_bnb_fp4_b200 import _lib, bnb_compat  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
code = torch.tensor(ext.BNB_FP4_CODE, device=dev)


def layer(N, K):
    return (torch.randint(0, 256, (N * K // 2, 1), dtype=torch.uint8, device=dev, generator=g),
            torch.rand(N * K // 64, device=dev, generator=g) * 0.1 + 0.01)


def check(y, x, p, a, N, K, what, tol=2e-2):
    w = ext.dequantize_fp4(p, a, 64, N, K, ext.float32)
    ref = x.float() @ w.t()
    err = float((y.float() - ref).abs().max() / ref.abs().max())
    print(f"{what:40s} err {err:.2e}", flush=True)
    assert err <= tol, what


for N, K, b, dt, st in ((256, 1024, 1, torch.bfloat16, ext.bfloat16), (320, 768, 3, torch.float16, ext.float16),
                        (160, 512, 8, torch.bfloat16, ext.bfloat16), (128, 256, 5, torch.float32, ext.float32)):
    p, a = layer(N, K)
    x = torch.randn(b, K, device=dev, generator=g).to(dt)
    check(ext.gemv_fp4(x, p, a, code, 64, st, [N, K]), x, p, a, N, K, f"stream gemv {N}x{K} b={b} {dt}")
    check(ext.gemv_fp4_bias(x, p, a, code, 64, st, [N, K], None, None, _lib.FLAG_FORCE_GENERIC), x, p, a, N, K,
          f"generic gemv {N}x{K} b={b}")
# nested
N, K = 256, 1024
p, a = layer(N, K)
code2 = bnb_compat.create_dynamic_map().to(dev)
off = float(a.mean())
q, am2 = bnb_compat.quantize_blockwise_8bit((a - off).cpu(), code2.cpu(), 256)
nd = ext.make_nested(q.to(dev), code2, am2.to(dev), off, 256)
a_dn = ext.absmax_denest(nd, N * K // 64, dev)
x = torch.randn(2, K, device=dev, generator=g).bfloat16()
check(ext.gemv_fp4_bias(x, p, None, code, 64, ext.bfloat16, [N, K], None, nd, 0), x, p, a_dn, N, K, "stream gemv nested")
# grouped, gated, residual
ps, as_ = zip(*[layer(n, 512) for n in (128, 64, 64)])
x = torch.randn(1, 512, device=dev, generator=g).bfloat16()
outs = ext.gemv_fp4_grouped(x, list(ps), list(as_), 64, ext.bfloat16, [[128, 512], [64, 512], [64, 512]])
for o, pp, aa, n in zip(outs, ps, as_, (128, 64, 64)):
    check(o, x, pp, aa, n, 512, f"grouped gemv member {n}")
pg, ag = layer(256, 512)
pu, au = layer(256, 512)
h = ext.gemv_fp4_fused(x, [pg, pu], [ag, au], 64, ext.bfloat16, [[256, 512], [256, 512]], gate_act="silu")[0]
wg, wu = ext.dequantize_fp4(pg, ag, 64, 256, 512, ext.float32), ext.dequantize_fp4(pu, au, 64, 256, 512, ext.float32)
ref = torch.nn.functional.silu(x.float() @ wg.t()) * (x.float() @ wu.t())
assert float((h.float() - ref).abs().max() / ref.abs().max()) <= 2e-2
res = torch.randn(1, 256, device=dev, generator=g).bfloat16()
y = ext.gemv_fp4_fused(x, [pg], [ag], 64, ext.bfloat16, [[256, 512]], residuals=[res])[0]
assert float((y.float() - (x.float() @ wg.t() + res.float())).abs().max()) <= 0.1 * float(ref.abs().max()) + 1.0
print("gated / residual epilogues ok", flush=True)
# tcgen05 GEMM
for M in (40, 300):
    N, K = 512, 1024
    p, a = layer(N, K)
    x = torch.randn(M, K, device=dev, generator=g).bfloat16()
    check(ext.gemm_fp4(x, p, a, code, N, K, 64), x, p, a, N, K, f"tcgen05 gemm M={M}", tol=2e-2)
# quantiser + dequant round trip
w = torch.randn(128, 256, device=dev, generator=g) * 0.05
pk, am = ext.quantize_fp4(w, 64)
d = ext.dequantize_fp4(pk, am, 64, 128, 256, ext.bfloat16)
assert float((d.float() - w).abs().max()) < 0.05
torch.cuda.synchronize()
print("sanitize_case: all kernels ran", flush=True)
