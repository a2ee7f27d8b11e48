"""Generate tests/golden/ref_*.npz by running the UNMODIFIED reference CUDA extension (built by
oracle/build_ref.py into oracle/_ref/) on a B200.  Run on the GPU box:

    gpurun -- 'python tests/golden/make_ref_golden.py gpurun_out/golden'

then copy gpurun_out/golden/ref_*.npz into tests/golden/ and commit them.  The files hold the seeded
inputs and the reference's outputs (dequantize_fp4 = tree kernel, dequantize_fp4_codebook, gemv_fp4)
for fp16/bf16/fp32; tests/test_oracle_cpu.py checks the CPU oracle against them, which is what pins
the oracle to the reference's behaviour.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402  (input synthesis only: the oracle quantiser)
from oracle.build_ref import load_module  # noqa: E402


def bits(t):
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.float32:
        return t.numpy()
    return t.view(torch.int16).numpy().view(np.uint16)


def case(ref, name, N, K, blocksize, seed, kind, outdir):
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(seed)
    n = N * K
    if kind == "gauss":
        w = (rng.standard_normal(n) * 0.02).astype(np.float32)
        packed, absmax = oracle.quantize(w, blocksize)
    else:  # every nibble value, arbitrary absmax (kept well inside the normal fp32 range)
        packed = rng.integers(0, 256, n // 2, dtype=np.uint8)
        absmax = (rng.random(n // blocksize) * 0.1 + 0.01).astype(np.float32)
    A = torch.from_numpy(packed).to(dev).view(-1, 1)
    am = torch.from_numpy(absmax).to(dev)
    code = torch.from_numpy(oracle.bnb_code()).to(dev)
    out = dict(n=n, N=N, K=K, blocksize=blocksize, packed=packed, absmax=absmax, seed=seed)
    for nm, dt, st in (("f16", torch.float16, ref.float16), ("bf16", torch.bfloat16, ref.bfloat16),
                       ("f32", torch.float32, ref.float32)):
        out[f"ref_tree_{nm}"] = bits(ref.dequantize_fp4(A, am, blocksize, N, K, st))
        out[f"ref_codebook_{nm}"] = bits(ref.dequantize_fp4_codebook(A, am, code, N, K, blocksize, n, st))
        x = torch.randn(1, K, generator=torch.Generator().manual_seed(seed + 100)).to(dt).to(dev)
        y = ref.gemv_fp4(x, A.t(), am, code, blocksize, st, [N, K])
        out[f"x_{nm}"] = bits(x).ravel()
        out[f"ref_gemv_{nm}"] = bits(y).ravel()
    torch.cuda.synchronize()
    np.savez_compressed(os.path.join(outdir, f"ref_{name}.npz"), **out)
    print("wrote", name, {k: getattr(v, "shape", v) for k, v in out.items() if k.startswith("ref_")})


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    ref = load_module()
    if ref is None:
        raise SystemExit("oracle/_ref/torch_bnb_fp4_ext_ref*.so not found: run oracle/build_ref.py first")
    case(ref, "gauss_64x1024", 64, 1024, 64, 11, "gauss", outdir)
    case(ref, "bytes_32x2048", 32, 2048, 64, 12, "bytes", outdir)
    case(ref, "gauss_16x4096_bs128", 16, 4096, 128, 13, "gauss", outdir)


if __name__ == "__main__":
    main()
