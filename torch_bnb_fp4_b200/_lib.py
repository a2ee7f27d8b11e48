"""ctypes binding of libfp4_b200.so (the C-ABI in include/fp4_b200.h).  There is no fallback: if the
library is missing or fails to load the import raises."""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# FP4_B200_LIB: a differently built copy of the library (kernel A/B experiments in tools/); default in-tree build
LIB_PATH = os.environ.get("FP4_B200_LIB") or os.path.join(HERE, "libfp4_b200.so")

F16, F32, BF16 = 0, 1, 2
FLAG_CODE_IS_BNB_FP4 = 1
FLAG_FORCE_GENERIC = 2
FLAG_NO_TMA = 4
FLAG_NO_I8 = 8
FLAG_NO_STREAM = 16

EXPORTS = [
    "fp4_b200_abi_version", "fp4_b200_status_string", "fp4_b200_dequantize",
    "fp4_b200_dequantize_nested", "fp4_b200_absmax_denest", "fp4_b200_gemv",
    "fp4_b200_gemv_workspace_bytes", "fp4_b200_gemv_grouped", "fp4_b200_gemv_grouped_tp", "fp4_b200_gemm",
    "fp4_b200_quantize", "fp4_b200_layer_create", "fp4_b200_layer_gemv", "fp4_b200_layer_destroy",
    "fp4_b200_layer_create_grouped", "fp4_b200_layer_gemv_grouped", "fp4_b200_launch_count",
    "fp4_b200_gemv_grouped_ex", "fp4_b200_layer_set_nested",
]


class Nested(ctypes.Structure):
    """fp4_b200_nested_t"""
    _fields_ = [("qabsmax", ctypes.c_void_p), ("code2", ctypes.c_void_p),
                ("absmax2", ctypes.c_void_p), ("offset", ctypes.c_float),
                ("blocksize2", ctypes.c_int)]


class TpExchange(ctypes.Structure):
    """fp4_b200_tp_t"""
    _fields_ = [("in_world", ctypes.c_int), ("in_base", ctypes.c_void_p), ("slot_bytes", ctypes.c_uint32),
                ("out_world", ctypes.c_int), ("out_rank", ctypes.c_int), ("out_peer_base", ctypes.c_void_p * 8),
                ("epochs", ctypes.c_void_p), ("err", ctypes.c_void_p)]


class Epilogue(ctypes.Structure):
    """fp4_b200_epilogue_t"""
    _fields_ = [("gate_act", ctypes.c_int), ("residual", ctypes.POINTER(ctypes.c_void_p)),
                ("nested", ctypes.POINTER(ctypes.POINTER(Nested)))]


GATE_ACT = {"silu": 1, "gelu_tanh": 2}


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m torch_bnb_fp4_b200.build` "
            "(there is no CPU or PyTorch fallback for the FP4 kernels)")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint
    lib.fp4_b200_abi_version.restype = i32
    lib.fp4_b200_status_string.restype = ctypes.c_char_p
    lib.fp4_b200_status_string.argtypes = [i32]
    lib.fp4_b200_dequantize.argtypes = [vp, vp, vp, vp, i64, i32, i32, vp]
    lib.fp4_b200_dequantize_nested.argtypes = [vp, ctypes.POINTER(Nested), vp, vp, i64, i32, i32, vp]
    lib.fp4_b200_absmax_denest.argtypes = [ctypes.POINTER(Nested), vp, i64, vp]
    lib.fp4_b200_gemv.argtypes = [vp, vp, vp, ctypes.POINTER(Nested), vp, vp, vp, i32, i32, i32,
                                  i32, i32, u32, vp, ctypes.c_size_t, vp]
    lib.fp4_b200_gemv_workspace_bytes.argtypes = [i32]
    lib.fp4_b200_layer_create.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, u32]
    lib.fp4_b200_layer_gemv.argtypes = [vp, vp, vp, i32, vp, ctypes.c_size_t, vp]
    lib.fp4_b200_layer_destroy.argtypes = [vp]
    lib.fp4_b200_layer_create_grouped.argtypes = [i32, ctypes.POINTER(vp), ctypes.POINTER(vp), vp, ctypes.POINTER(vp),
                                                  ctypes.POINTER(i32), i32, i32, i32, u32]
    lib.fp4_b200_layer_set_nested.argtypes = [vp, i32, ctypes.POINTER(Nested)]
    lib.fp4_b200_layer_gemv_grouped.argtypes = [vp, vp, ctypes.POINTER(vp), i32, ctypes.POINTER(TpExchange), vp]
    lib.fp4_b200_gemv_grouped.argtypes = [vp, i32, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp),
                                          ctypes.POINTER(vp), ctypes.POINTER(i32), i32, i32, i32, i32, u32, vp]
    lib.fp4_b200_gemv_grouped_tp.argtypes = [vp, i32, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp),
                                             ctypes.POINTER(vp), ctypes.POINTER(i32), i32, i32, i32, i32, u32,
                                             ctypes.POINTER(TpExchange), vp]
    lib.fp4_b200_gemv_grouped_ex.argtypes = [vp, i32, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp),
                                             ctypes.POINTER(vp), ctypes.POINTER(i32), i32, i32, i32, i32, u32,
                                             ctypes.POINTER(TpExchange), ctypes.POINTER(Epilogue), vp]
    lib.fp4_b200_gemm.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, u32, vp,
                                  ctypes.c_size_t, vp]
    lib.fp4_b200_quantize.argtypes = [vp, i32, i64, i32, vp, vp, vp]
    for name in EXPORTS:
        getattr(lib, name)  # AttributeError if the ABI is incomplete
        if name not in ("fp4_b200_status_string", "fp4_b200_gemv_workspace_bytes", "fp4_b200_layer_create",
                        "fp4_b200_layer_create_grouped", "fp4_b200_layer_destroy", "fp4_b200_launch_count"):
            getattr(lib, name).restype = i32
    lib.fp4_b200_gemv_workspace_bytes.restype = ctypes.c_size_t
    lib.fp4_b200_layer_create.restype = vp
    lib.fp4_b200_launch_count.restype = ctypes.c_ulonglong
    lib.fp4_b200_layer_create_grouped.restype = vp
    lib.fp4_b200_layer_destroy.restype = None
    if lib.fp4_b200_abi_version() != 1:
        raise ImportError("libfp4_b200.so ABI version mismatch")
    return lib


lib = _load()


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib.fp4_b200_status_string(status).decode()
        raise RuntimeError(f"{what}: {msg} (status {status})")
