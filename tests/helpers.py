"""Shared helpers for the parity tests (numpy <-> torch plumbing, seeded synthetic inputs)."""
import numpy as np
import torch

import oracle

NP2O = {torch.float16: oracle.F16, torch.bfloat16: oracle.BF16, torch.float32: oracle.F32}
DTYPES = [torch.float16, torch.bfloat16, torch.float32]


def synth_quant(n, blocksize=64, seed=0, scale=0.02):
    """Seeded gaussian weights -> oracle quantiser -> (packed u8, absmax f32, w f32)."""
    rng = np.random.default_rng(seed)
    w = (rng.standard_normal(n) * scale).astype(np.float32)
    packed, absmax = oracle.quantize(w, blocksize)
    return packed, absmax, w


def synth_bytes(n, blocksize=64, seed=0):
    """Uniform random nibbles + random positive absmax (exercises every code incl. -0)."""
    rng = np.random.default_rng(seed)
    packed = rng.integers(0, 256, (n + 1) // 2, dtype=np.uint8)
    absmax = (rng.random((n + blocksize - 1) // blocksize) * 0.1 + 0.01).astype(np.float32)
    return packed, absmax


def bits_of(t: torch.Tensor) -> np.ndarray:
    """Bit patterns of a CUDA/CPU tensor as uint16 (16-bit dtypes) or uint32 (fp32)."""
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.float32:
        return t.view(torch.int32).numpy().view(np.uint32).ravel()
    return t.view(torch.int16).numpy().view(np.uint16).ravel()


def oracle_bits(a: np.ndarray) -> np.ndarray:
    return a.view(np.uint32).ravel() if a.dtype == np.float32 else a.ravel()


def to_dev(a: np.ndarray, dev, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return t if dtype is None else t.to(dtype)


def round_x(x32: np.ndarray, dtype: torch.dtype) -> np.ndarray:
    """fp32 values rounded to `dtype` (as fp32 values) — what the kernel actually sees."""
    return torch.from_numpy(x32).to(dtype).float().numpy()


def normwise(a: np.ndarray, ref: np.ndarray) -> float:
    return float(np.max(np.abs(a.astype(np.float64) - ref.astype(np.float64))) /
                 max(np.max(np.abs(ref.astype(np.float64))), 1e-30))
