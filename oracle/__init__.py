"""CPU oracle for the FP4 Linear hot path — TEST INFRASTRUCTURE ONLY.

Thin ctypes/numpy wrapper over ``oracle/fp4_oracle.c`` (the restatement, with reference file:line
citations).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package; the product path never does.

Parity pin: see the header of ``fp4_oracle.c`` and ``tests/golden/README.md``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfp4_oracle.so")

F16, F32, BF16 = 0, 1, 2
_NP_OUT = {F16: np.uint16, BF16: np.uint16, F32: np.float32}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "fp4_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-o", _SO, src, "-lm"]
        )
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.fp4o_f32_to_f16.restype = ctypes.c_uint16
        _lib.fp4o_f32_to_f16.argtypes = [ctypes.c_float]
        _lib.fp4o_f32_to_bf16.restype = ctypes.c_uint16
        _lib.fp4o_f32_to_bf16.argtypes = [ctypes.c_float]
        _lib.fp4o_num_threads.restype = ctypes.c_int
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


def bnb_code() -> np.ndarray:
    out = np.empty(16, np.float32)
    lib().fp4o_bnb_code(_p(out))
    return out


def ref_code_param() -> np.ndarray:
    out = np.empty(16, np.float32)
    lib().fp4o_ref_code_param(_p(out))
    return out


def num_threads() -> int:
    return int(lib().fp4o_num_threads())


def f32_to_bits(x: np.ndarray, dtype: int) -> np.ndarray:
    """Round fp32 values to fp16/bf16 bit patterns with the oracle's scalar converters."""
    x = _c(x, np.float32).ravel()
    fn = lib().fp4o_f32_to_f16 if dtype == F16 else lib().fp4o_f32_to_bf16
    return np.array([fn(float(v)) for v in x], dtype=np.uint16)


def bits_to_f32(bits: np.ndarray, dtype: int) -> np.ndarray:
    """fp16/bf16 bit patterns (uint16) or fp32 values -> fp32 values (numpy only)."""
    if dtype == F32:
        return np.asarray(bits, np.float32)
    b = np.asarray(bits, np.uint16)
    if dtype == F16:
        return b.view(np.float16).astype(np.float32)
    return (b.astype(np.uint32) << 16).view(np.float32)


def quantize(w: np.ndarray, blocksize: int = 64):
    w = _c(w, np.float32).ravel()
    n = w.size
    packed = np.zeros((n + 1) // 2, np.uint8)
    absmax = np.zeros((n + blocksize - 1) // blocksize, np.float32)
    lib().fp4o_quantize(_p(w), ctypes.c_int64(n), ctypes.c_int(blocksize), _p(packed), _p(absmax))
    return packed, absmax


def dequant_tree(packed, absmax, n: int, blocksize: int, dtype: int) -> np.ndarray:
    """Returns fp32 values for F32, uint16 bit patterns for F16/BF16."""
    packed, absmax = _c(packed, np.uint8).ravel(), _c(absmax, np.float32).ravel()
    out = np.empty(n, _NP_OUT[dtype])
    lib().fp4o_dequant_tree(_p(packed), _p(absmax), ctypes.c_int64(n), ctypes.c_int(blocksize),
                            ctypes.c_int(dtype), _p(out))
    return out


def dequant_code(packed, absmax, code16, n: int, blocksize: int, dtype: int) -> np.ndarray:
    packed, absmax = _c(packed, np.uint8).ravel(), _c(absmax, np.float32).ravel()
    code16 = _c(code16, np.float32).ravel()
    assert code16.size == 16
    out = np.empty(n, _NP_OUT[dtype])
    lib().fp4o_dequant_code(_p(packed), _p(absmax), _p(code16), ctypes.c_int64(n),
                            ctypes.c_int(blocksize), ctypes.c_int(dtype), _p(out))
    return out


def denest(qabsmax, code2, absmax2, offset: float, blocksize2: int) -> np.ndarray:
    qabsmax = _c(qabsmax, np.uint8).ravel()
    code2, absmax2 = _c(code2, np.float32).ravel(), _c(absmax2, np.float32).ravel()
    out = np.empty(qabsmax.size, np.float32)
    lib().fp4o_denest(_p(qabsmax), _p(code2), _p(absmax2), ctypes.c_float(offset),
                      ctypes.c_int(blocksize2), ctypes.c_int64(qabsmax.size), _p(out))
    return out


def linear_f64(x, packed, absmax, code16, bias, N: int, K: int, blocksize: int,
               wdtype: int = F32) -> np.ndarray:
    """x: [batch, K] fp32 values -> [batch, N] float64."""
    x = _c(x, np.float32).reshape(-1, K)
    batch = x.shape[0]
    packed, absmax = _c(packed, np.uint8).ravel(), _c(absmax, np.float32).ravel()
    code16 = _c(code16, np.float32).ravel()
    b = None if bias is None else _c(bias, np.float32).ravel()
    out = np.empty((batch, N), np.float64)
    lib().fp4o_linear_f64(_p(x), _p(packed), _p(absmax), _p(code16),
                          _p(b) if b is not None else None, ctypes.c_int(batch), ctypes.c_int(N),
                          ctypes.c_int(K), ctypes.c_int(blocksize), ctypes.c_int(wdtype), _p(out))
    return out


def linear_f32(x, packed, absmax, code16, N: int, K: int, blocksize: int) -> np.ndarray:
    """The timed CPU baseline: fp32 dequant + fp32 dot, OpenMP over rows."""
    x = _c(x, np.float32).reshape(-1, K)
    batch = x.shape[0]
    packed, absmax = _c(packed, np.uint8).ravel(), _c(absmax, np.float32).ravel()
    code16 = _c(code16, np.float32).ravel()
    out = np.empty((batch, N), np.float32)
    lib().fp4o_linear_f32(_p(x), _p(packed), _p(absmax), _p(code16), ctypes.c_int(batch),
                          ctypes.c_int(N), ctypes.c_int(K), ctypes.c_int(blocksize), _p(out))
    return out


def gemv_ref_emulate(x, packed, absmax, N: int, K: int, blocksize: int, dtype: int) -> np.ndarray:
    """Emulates the reference GEMV kernel's arithmetic (batch 1). x: fp32 values rounded to dtype."""
    x = _c(x, np.float32).ravel()
    assert x.size == K and K % 32 == 0
    packed, absmax = _c(packed, np.uint8).ravel(), _c(absmax, np.float32).ravel()
    out = np.empty(N, np.float32)
    lib().fp4o_gemv_ref_emulate(_p(x), _p(packed), _p(absmax), ctypes.c_int(N), ctypes.c_int(K),
                                ctypes.c_int(blocksize), ctypes.c_int(dtype), _p(out))
    return out
