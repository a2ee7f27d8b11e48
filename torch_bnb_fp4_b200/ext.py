"""``torch_bnb_fp4_ext`` for B200: the reference's C++-extension op surface, same names, same
positional signatures, same return shapes (reference csrc/torch_fp4.cpp:125-139), implemented as a
thin binding over the C-ABI library ``libfp4_b200.so`` (include/fp4_b200.h).

Differences from the reference binding, all deliberate:
  * kernels launch on the CURRENT stream of the tensors' device (the reference uses the legacy
    default stream with no device guard, csrc/gemv_fp4_optimized.cu:266), so every op is CUDA-graph
    capturable; nothing synchronises;
  * errors raise (RuntimeError / TypeError) instead of printing (csrc/dequant_fp4_optimized.cu:48-53,201-203);
  * ``dequantize_fp4_codebook`` and ``gemv_fp4`` honour the ``code`` tensor they are handed (the
    reference ignores it and uses its hard-coded CODE_PARAM, csrc/dequant_fp4_optimized.cu:207-255);
  * ``gemv_fp4`` computes every row of A for batch 1..8 (the reference fills row 0 only, n=1 at
    csrc/gemv_fp4_optimized.cu:289), accumulates in fp32, and ``gemv_fp4_bias`` fuses the bias;
  * ``qlinear_codebook*`` dequantise the whole weight (the reference passes the byte count as the
    element count and leaves half the matrix uninitialised, csrc/torch_fp4.cpp:90,101 — SURVEY N3).
There is no CPU path: CPU tensors raise, exactly like the reference's CHECK_CUDA.
"""
from __future__ import annotations

import ctypes
import enum
import weakref
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import Nested, check, lib


class ScalarType(enum.IntEnum):
    """Mirror of the pybind enum (reference csrc/torch_fp4.cpp:22-26,126-130); values are the
    C-ABI dtype codes."""
    float16 = _lib.F16
    float32 = _lib.F32
    bfloat16 = _lib.BF16


# .export_values() in the reference: the members are also module attributes
float16 = ScalarType.float16
float32 = ScalarType.float32
bfloat16 = ScalarType.bfloat16

_TORCH_DTYPE = {ScalarType.float16: torch.float16, ScalarType.float32: torch.float32,
                ScalarType.bfloat16: torch.bfloat16}
_CODE_OF = {torch.float16: _lib.F16, torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}

BNB_FP4_CODE = (0.0, 5.208333333e-03, 0.66666667, 1.0, 0.33333333, 0.5, 0.16666667, 0.25,
                -0.0, -5.208333333e-03, -0.66666667, -1.0, -0.33333333, -0.5, -0.16666667, -0.25)


def get_scalar_type(t) -> torch.dtype:
    """csrc/torch_fp4.cpp:28-39: bad enum -> TypeError."""
    try:
        return _TORCH_DTYPE[ScalarType(t)]
    except (ValueError, KeyError, TypeError):
        raise TypeError("Unsupported scalar type") from None


def _check_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")


def _check_contig(t: torch.Tensor, name: str) -> None:
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def _check_in(t: torch.Tensor, name: str, dtype: Optional[torch.dtype] = None) -> None:
    _check_cuda(t, name)
    _check_contig(t, name)
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")


class _on_device:
    """Device guard + current stream of the tensor's device (c10::cuda::CUDAGuard equivalent)."""
    __slots__ = ("idx", "prev")

    def __init__(self, t: torch.Tensor):
        self.idx = t.device.index

    def __enter__(self) -> int:
        self.prev = torch.cuda.current_device()
        if self.prev != self.idx:
            torch.cuda.set_device(self.idx)
        return torch.cuda.current_stream(self.idx).cuda_stream

    def __exit__(self, *exc):
        if self.prev != self.idx:
            torch.cuda.set_device(self.prev)
        return False


# ---- codebook identity cache ---------------------------------------------------------------------
# The integer tensor-core GEMV is only valid for the bitsandbytes FP4 table.  Whether a given `code`
# tensor holds it is checked ONCE per tensor object (one device->host read, outside the hot loop and
# outside graph capture) and remembered while the tensor is alive and unmodified.
_code_cache: dict = {}


def code_is_bnb_fp4(code: torch.Tensor) -> bool:
    key = id(code)
    hit = _code_cache.get(key)
    if hit is not None and hit[0]() is code and hit[1] == code._version:
        return hit[2]
    if code.is_cuda and torch.cuda.is_current_stream_capturing():
        return False  # cannot read the table during capture: take the generic kernel
    # compared by VALUE: bitsandbytes builds QuantState.code from a Python list whose entry 8 is the int literal
    # -0, i.e. +0.0, while the reference's table has -0.0 there; the sign of a zero weight changes no sum
    ref = torch.tensor(BNB_FP4_CODE, dtype=torch.float32)
    ok = bool(code.numel() == 16 and code.dtype == torch.float32
              and torch.equal(code.detach().reshape(-1).cpu(), ref))
    _code_cache[key] = (weakref.ref(code, lambda _r, k=key: _code_cache.pop(k, None)),
                        code._version, ok)
    return ok


# ---- GEMV workspace --------------------------------------------------------------------------------
# ABI version 1 kernels parked partial sums in a caller-provided scratch buffer; none of today's kernels needs one
# (fp4_b200_gemv_workspace_bytes returns 0), so NULL is passed.
def _gemv_workspace(device: torch.device, stream_ptr: int, n_out: int):
    return None


def make_nested(qabsmax: torch.Tensor, code2: torch.Tensor, absmax2: torch.Tensor, offset: float,
                blocksize2: int) -> Nested:
    """Pack a bitsandbytes nested state (state2 + offset) for the *_nested entry points."""
    _check_in(qabsmax, "qabsmax", torch.uint8)
    _check_in(code2, "code2", torch.float32)
    _check_in(absmax2, "absmax2", torch.float32)
    if code2.numel() != 256:
        raise RuntimeError("nested code2 must have 256 entries")
    n = Nested(qabsmax.data_ptr(), code2.data_ptr(), absmax2.data_ptr(), float(offset),
               int(blocksize2))
    n._keep = (qabsmax, code2, absmax2)  # keep the tensors alive as long as the struct
    return n


# ---- reference op surface --------------------------------------------------------------------------
def dequantize_fp4(A: torch.Tensor, absmax: torch.Tensor, blocksize: int, M: int, N: int,
                   o_type) -> torch.Tensor:
    """reference csrc/torch_fp4.cpp:41-50 (tree decoder == bitsandbytes constants)."""
    _check_in(A, "A", torch.uint8)
    _check_in(absmax, "absmax", torch.float32)
    dt = get_scalar_type(o_type)
    out = torch.empty((M, N), dtype=dt, device=A.device)
    n = M * N
    if A.numel() * 2 < n:
        raise RuntimeError(f"A holds {A.numel()} bytes, fewer than M*N/2 = {n / 2}")
    with _on_device(A) as st:
        check(lib.fp4_b200_dequantize(A.data_ptr(), absmax.data_ptr(), None, out.data_ptr(), n,
                                      blocksize, _CODE_OF[dt], st), "dequantize_fp4")
    return out


def dequantize_fp4_codebook(A: torch.Tensor, absmax: torch.Tensor, codebook: torch.Tensor, M: int,
                            N: int, blocksize: int, n: int, dtype) -> torch.Tensor:
    """reference csrc/torch_fp4.cpp:52-62; `n` = number of ELEMENTS to dequantise (the tail of the
    [M, N] output beyond n is left uninitialised, as in the reference)."""
    _check_in(A, "A", torch.uint8)
    _check_in(absmax, "absmax", torch.float32)
    _check_in(codebook, "codebook", torch.float32)
    if codebook.numel() != 16:
        raise RuntimeError("codebook must have 16 entries")
    dt = get_scalar_type(dtype)
    if n < 0 or n > M * N or A.numel() * 2 < n:
        raise RuntimeError(f"n={n} inconsistent with M*N={M * N} and {A.numel()} packed bytes")
    out = torch.empty((M, N), dtype=dt, device=A.device)
    with _on_device(A) as st:
        check(lib.fp4_b200_dequantize(A.data_ptr(), absmax.data_ptr(), codebook.data_ptr(),
                                      out.data_ptr(), n, blocksize, _CODE_OF[dt], st),
              "dequantize_fp4_codebook")
    return out


def dequantize_fp4_nested(A: torch.Tensor, nested: Nested, codebook: Optional[torch.Tensor], M: int,
                          N: int, blocksize: int, dtype) -> torch.Tensor:
    """Extension: dequantise with a double-quantised absmax decoded in the kernel."""
    _check_in(A, "A", torch.uint8)
    dt = get_scalar_type(dtype)
    out = torch.empty((M, N), dtype=dt, device=A.device)
    with _on_device(A) as st:
        check(lib.fp4_b200_dequantize_nested(
            A.data_ptr(), ctypes.byref(nested), None if codebook is None else codebook.data_ptr(),
            out.data_ptr(), M * N, blocksize, _CODE_OF[dt], st), "dequantize_fp4_nested")
    return out


def absmax_denest(nested: Nested, nblocks: int, device) -> torch.Tensor:
    out = torch.empty(nblocks, dtype=torch.float32, device=device)
    with _on_device(out) as st:
        check(lib.fp4_b200_absmax_denest(ctypes.byref(nested), out.data_ptr(), nblocks, st),
              "absmax_denest")
    return out


def _gemv(A, B, absmax, datatype, blocksize, dtype, Bshape, bias, nested, flags, what):
    _check_in(A, "A")
    _check_in(B, "B", torch.uint8)
    if nested is None:
        _check_in(absmax, "absmax", torch.float32)
    if datatype is not None:
        _check_in(datatype, "datatype", torch.float32)
    dt = get_scalar_type(dtype)
    if A.dtype != dt:
        raise RuntimeError(f"A is {A.dtype} but dtype argument says {dt}")
    n_out, k = int(Bshape[0]), int(Bshape[1])
    if A.dim() < 1 or A.shape[-1] != k:
        raise RuntimeError(f"A must be [..., {k}], got {tuple(A.shape)}")
    batch = A.numel() // k if k else 0
    if B.numel() * 2 < n_out * k:
        raise RuntimeError("B holds fewer than N*K/2 bytes")
    if bias is not None:
        _check_in(bias, "bias", dt)
        if bias.numel() != n_out:
            raise RuntimeError("bias must have N entries")
    out = torch.empty(A.shape[:-1] + (n_out,), dtype=dt, device=A.device)
    if batch == 0 or n_out == 0:
        return out
    if datatype is not None and code_is_bnb_fp4(datatype):
        flags |= _lib.FLAG_CODE_IS_BNB_FP4
    with _on_device(A) as st:
        check(lib.fp4_b200_gemv(
            A.data_ptr(), B.data_ptr(), None if absmax is None else absmax.data_ptr(),
            None if nested is None else ctypes.byref(nested),
            None if datatype is None else datatype.data_ptr(),
            None if bias is None else bias.data_ptr(), out.data_ptr(), batch, n_out, k, blocksize,
            _CODE_OF[dt], flags, None, 0, st), what)
    return out


class GemvLauncher:
    """Pre-validated launcher for ONE layer's decode GEMV (the module's hot path): the checks of `_gemv` are
    done once here, the per-call work is one torch.empty, one raw-stream query and one ctypes call.  The
    launcher keeps the tensors alive, so the cached device pointers stay valid; anything it was not built
    for (another device current, non-contiguous input, another dtype) is the caller's slow path."""
    __slots__ = ("keep", "handle", "n_out", "k", "dt", "dev", "idx", "what", "_shapes")

    def __init__(self, B, absmax, datatype, blocksize, dtype, Bshape, bias):
        _check_in(B, "B", torch.uint8)
        _check_in(absmax, "absmax", torch.float32)
        self.dt = get_scalar_type(dtype)
        self.n_out, self.k = int(Bshape[0]), int(Bshape[1])
        if B.numel() * 2 < self.n_out * self.k:
            raise RuntimeError("B holds fewer than N*K/2 bytes")
        flags = 0
        pcode = None
        if datatype is not None:
            _check_in(datatype, "datatype", torch.float32)
            pcode = datatype.data_ptr()
            if code_is_bnb_fp4(datatype):
                flags |= _lib.FLAG_CODE_IS_BNB_FP4
        pbias = None
        if bias is not None:
            _check_in(bias, "bias", self.dt)
            if bias.numel() != self.n_out:
                raise RuntimeError("bias must have N entries")
            pbias = bias.data_ptr()
        self.keep = (B, absmax, datatype, bias)
        # the constants live in a prepared-layer handle of the C-ABI: 7 marshalled arguments per call, not 16
        self.handle = lib.fp4_b200_layer_create(B.data_ptr(), absmax.data_ptr(), pcode, pbias, self.n_out, self.k,
                                                int(blocksize), _CODE_OF[self.dt], flags)
        if not self.handle:
            raise RuntimeError("fp4_b200_layer_create failed")
        self.dev, self.idx = B.device, B.device.index
        self.what = "gemv_fp4_bias"
        self._shapes = {}

    def __del__(self):
        try:
            h, self.handle = getattr(self, "handle", None), None
            if h and lib is not None:
                lib.fp4_b200_layer_destroy(h)
        except Exception:  # noqa: BLE001 - interpreter shutdown: the library may already be gone
            pass

    def __call__(self, A: torch.Tensor, batch: int) -> torch.Tensor:
        """A: contiguous CUDA tensor [..., K] of the launcher's dtype on the launcher's (current) device."""
        shp = self._shapes.get(A.shape)
        if shp is None:
            shp = self._shapes[A.shape] = tuple(A.shape[:-1]) + (self.n_out,)
        out = torch.empty(shp, dtype=self.dt, device=self.dev)
        st = torch._C._cuda_getCurrentRawStream(self.idx)
        rc = lib.fp4_b200_layer_gemv(self.handle, A.data_ptr(), out.data_ptr(), batch, None, 0, st)
        if rc:
            check(rc, self.what)
        return out


def gemv_fp4(A: torch.Tensor, B: torch.Tensor, absmax: torch.Tensor, datatype: torch.Tensor,
             blocksize: int, dtype, Bshape: Sequence[int]) -> torch.Tensor:
    """reference csrc/torch_fp4.cpp:105-123 -> csrc/gemv_fp4_optimized.cu:277-368.
    A: [batch, K] (or [b0, b1, K]) with batch <= 8; returns [batch, N] / [b0, b1, N]."""
    return _gemv(A, B, absmax, datatype, blocksize, dtype, Bshape, None, None, 0, "gemv_fp4")


def gemv_fp4_bias(A, B, absmax, datatype, blocksize, dtype, Bshape, bias=None, nested=None,
                  flags: int = 0) -> torch.Tensor:
    """Extension: bias fused in the epilogue (replaces the separate `out += bias`,
    reference torch_bnb_fp4/__init__.py:608-613) and optional in-kernel nested absmax."""
    return _gemv(A, B, absmax, datatype, blocksize, dtype, Bshape, bias, nested, flags,
                 "gemv_fp4_bias")


def gemv_fp4_grouped(A: Optional[torch.Tensor], Bs: Sequence[torch.Tensor], absmaxes: Sequence[torch.Tensor],
                     blocksize: int, dtype, Bshapes: Sequence[Sequence[int]],
                     biases: Optional[Sequence[Optional[torch.Tensor]]] = None, tp=None,
                     outs: Optional[Sequence[torch.Tensor]] = None, batch_shape: Optional[Sequence[int]] = None):
    """Extension: ONE launch for several bitsandbytes-FP4 weights that share the input (q/k/v, gate/up).
    Returns a list of outputs equal (up to fp32 summation order) to calling gemv_fp4_bias per weight, or None when the shapes are
    outside the grouped kernel's domain (the caller then issues the calls one by one).
    `tp` (a _lib.TpExchange, see include/fp4_b200.h) switches on the tensor-parallel exchange through peer
    memory: with tp.in_world > 1 A may be None (x is the sum of the ranks' partials; give batch_shape),
    with tp.out_world > 1 `outs` must hold the base of this rank's exchange buffer."""
    dt = get_scalar_type(dtype)
    n = len(Bs)
    if n < 1 or n > 4 or len(absmaxes) != n or len(Bshapes) != n:
        return None
    k = int(Bshapes[0][1])
    if A is not None:
        _check_in(A, "A")
        if A.dtype != dt:
            raise RuntimeError(f"A is {A.dtype} but dtype argument says {dt}")
        if A.shape[-1] != k:
            raise RuntimeError("grouped GEMV: every weight must have in_features == A.shape[-1]")
        lead = tuple(A.shape[:-1])
        batch = A.numel() // k if k else 0
    else:
        if tp is None or tp.in_world <= 1 or batch_shape is None:
            raise RuntimeError("grouped GEMV without an input needs a tensor-parallel exchange and batch_shape")
        lead = tuple(batch_shape)
        batch = 1
        for v in lead:
            batch *= int(v)
    if any(int(sh[1]) != k for sh in Bshapes):
        raise RuntimeError("grouped GEMV: every weight must have the same in_features")
    if batch < 1 or batch > 8:
        return None
    for B, am in zip(Bs, absmaxes):
        _check_in(B, "B", torch.uint8)
        _check_in(am, "absmax", torch.float32)
    dev = Bs[0].device
    if outs is None:
        outs = [torch.empty(lead + (int(sh[0]),), dtype=dt, device=dev) for sh in Bshapes]
    vp = ctypes.c_void_p
    pk = (vp * n)(*[B.data_ptr() for B in Bs])
    am = (vp * n)(*[a.data_ptr() for a in absmaxes])
    ou = (vp * n)(*[o.data_ptr() for o in outs])
    ns = (ctypes.c_int * n)(*[int(sh[0]) for sh in Bshapes])
    bi = None
    if biases is not None and any(b is not None for b in biases):
        for b in biases:
            if b is not None:
                _check_in(b, "bias", dt)
        bi = (vp * n)(*[None if b is None else b.data_ptr() for b in biases])
    with _on_device(Bs[0]) as st:
        status = lib.fp4_b200_gemv_grouped_tp(None if A is None else A.data_ptr(), n, pk, am, bi, ou, ns, batch, k,
                                              blocksize, _CODE_OF[dt], _lib.FLAG_CODE_IS_BNB_FP4,
                                              None if tp is None else ctypes.byref(tp), st)
    if status == -7:  # FP4_B200_ERR_UNSUPPORTED
        return None
    check(status, "gemv_fp4_grouped")
    return list(outs)


def gemv_fp4_fused(A: torch.Tensor, Bs: Sequence[torch.Tensor], absmaxes: Sequence[torch.Tensor], blocksize: int, dtype,
                   Bshapes: Sequence[Sequence[int]], biases: Optional[Sequence[Optional[torch.Tensor]]] = None,
                   gate_act: Optional[str] = None, residuals: Optional[Sequence[Optional[torch.Tensor]]] = None):
    """Extension: the grouped fused dequant-GEMV with the neighbouring elementwise ops in its
    epilogue.  gate_act = "silu" / "gelu_tanh": Bs = (gate, up) of a gated MLP, returns ONE tensor
    act(x W_gate^T + b_gate) * (x W_up^T + b_up) computed in fp32 before the single rounding to the output dtype.
    residuals: tensors [..., N_m] added to the outputs (the residual stream around o / down).  Returns a list of
    outputs, or None when the shapes are outside the streaming kernel (the caller composes the ops itself)."""
    dt = get_scalar_type(dtype)
    n = len(Bs)
    k = int(Bshapes[0][1])
    _check_in(A, "A")
    if A.dtype != dt or A.shape[-1] != k or any(int(sh[1]) != k for sh in Bshapes):
        raise RuntimeError("fused GEMV: A / weights disagree on dtype or in_features")
    batch = A.numel() // k if k else 0
    if n < 1 or n > 4 or batch < 1 or batch > 8:
        return None
    for B, am in zip(Bs, absmaxes):
        _check_in(B, "B", torch.uint8)
        _check_in(am, "absmax", torch.float32)
    lead = tuple(A.shape[:-1])
    vp = ctypes.c_void_p
    epi = _lib.Epilogue()
    n_out = 1 if gate_act else n
    if gate_act:
        if gate_act not in _lib.GATE_ACT or n != 2 or int(Bshapes[0][0]) != int(Bshapes[1][0]):
            raise RuntimeError("gate_act needs the gate and up projections of one MLP (same out_features)")
        epi.gate_act = _lib.GATE_ACT[gate_act]
    outs = [torch.empty(lead + (int(Bshapes[i][0]),), dtype=dt, device=A.device) for i in range(n_out)]
    keep = None
    if residuals is not None and any(r is not None for r in residuals):
        if gate_act:
            raise RuntimeError("a residual goes with a plain projection, not with the gated pair")
        for r, o in zip(residuals, outs):
            if r is not None:
                _check_in(r, "residual", dt)
                if r.shape != o.shape:
                    raise RuntimeError("residual must have the shape of the output")
        keep = (vp * n)(*[None if r is None else r.data_ptr() for r in residuals])
        epi.residual = ctypes.cast(keep, ctypes.POINTER(vp))
    pk = (vp * n)(*[B.data_ptr() for B in Bs])
    am = (vp * n)(*[a.data_ptr() for a in absmaxes])
    ou = (vp * n)(*([o.data_ptr() for o in outs] + [None] * (n - n_out)))
    ns = (ctypes.c_int * n)(*[int(sh[0]) for sh in Bshapes])
    bi = None
    if biases is not None and any(b is not None for b in biases):
        for b in biases:
            if b is not None:
                _check_in(b, "bias", dt)
        bi = (vp * n)(*[None if b is None else b.data_ptr() for b in biases])
    with _on_device(A) as st:
        status = lib.fp4_b200_gemv_grouped_ex(A.data_ptr(), n, pk, am, bi, ou, ns, batch, k, blocksize, _CODE_OF[dt],
                                              _lib.FLAG_CODE_IS_BNB_FP4, None, ctypes.byref(epi), st)
    if status == -7:
        return None
    check(status, "gemv_fp4_fused")
    return outs


class GroupLauncher:
    """Pre-validated launcher for a group of layers that share the input (see GemvLauncher): the pointer
    arrays of fp4_b200_gemv_grouped are built once, a call allocates the outputs and fills in their pointers."""
    __slots__ = ("keep", "n", "pk", "am", "bi", "ou", "ns", "n_outs", "k", "blocksize", "dt", "dtcode", "dev",
                 "idx", "unsupported", "handle")

    def __init__(self, Bs, absmaxes, blocksize, dtype, Bshapes, biases=None, nesteds=None):
        """absmaxes[i] may be None when nesteds[i] (a _lib.Nested from make_nested) is given: that member keeps its
        absmax double-quantised and the kernel decodes it."""
        self.dt = get_scalar_type(dtype)
        n = self.n = len(Bs)
        if n < 1 or n > 4 or len(absmaxes) != n or len(Bshapes) != n:
            raise RuntimeError("a group holds 1..4 layers")
        self.k = int(Bshapes[0][1])
        if any(int(sh[1]) != self.k for sh in Bshapes):
            raise RuntimeError("grouped GEMV: every weight must have the same in_features")
        nesteds = list(nesteds) if nesteds is not None else [None] * n
        for B, a, nd in zip(Bs, absmaxes, nesteds):
            _check_in(B, "B", torch.uint8)
            if nd is None:
                _check_in(a, "absmax", torch.float32)
        vp = ctypes.c_void_p
        self.n_outs = [int(sh[0]) for sh in Bshapes]
        self.pk = (vp * n)(*[B.data_ptr() for B in Bs])
        self.am = (vp * n)(*[None if a is None else a.data_ptr() for a in absmaxes])
        self.ou = (vp * n)()
        self.ns = (ctypes.c_int * n)(*self.n_outs)
        self.bi = None
        if biases is not None and any(b is not None for b in biases):
            for b in biases:
                if b is not None:
                    _check_in(b, "bias", self.dt)
            self.bi = (vp * n)(*[None if b is None else b.data_ptr() for b in biases])
        self.keep = (list(Bs), list(absmaxes), None if biases is None else list(biases), nesteds)
        self.blocksize, self.dtcode = int(blocksize), _CODE_OF[self.dt]
        self.dev, self.idx = Bs[0].device, Bs[0].device.index
        self.unsupported = set()  # batch sizes the grouped kernel refused (FP4_B200_ERR_UNSUPPORTED)
        # prepared group handle of the C-ABI: the constants are bound once
        self.handle = lib.fp4_b200_layer_create_grouped(n, self.pk, self.am, None, self.bi, self.ns, self.k,
                                                        self.blocksize, self.dtcode, _lib.FLAG_CODE_IS_BNB_FP4)
        if not self.handle:
            raise RuntimeError("fp4_b200_layer_create_grouped failed")
        for i, nd in enumerate(nesteds):
            if nd is not None:
                check(lib.fp4_b200_layer_set_nested(self.handle, i, ctypes.byref(nd)), "fp4_b200_layer_set_nested")

    def __del__(self):
        try:
            h, self.handle = getattr(self, "handle", None), None
            if h and lib is not None:
                lib.fp4_b200_layer_destroy(h)
        except Exception:  # noqa: BLE001 - interpreter shutdown: the library may already be gone
            pass

    def __call__(self, A: torch.Tensor, batch: int):
        """A: contiguous [..., K] of the launcher's dtype on the launcher's (current) device; None = unsupported."""
        if batch in self.unsupported:
            return None
        lead = tuple(A.shape[:-1])
        outs = [torch.empty(lead + (m,), dtype=self.dt, device=self.dev) for m in self.n_outs]
        for i, o in enumerate(outs):
            self.ou[i] = o.data_ptr()
        st = torch._C._cuda_getCurrentRawStream(self.idx)
        rc = lib.fp4_b200_layer_gemv_grouped(self.handle, A.data_ptr(), self.ou, batch, None, st)
        if rc == -7:
            self.unsupported.add(batch)
            return None
        if rc:
            check(rc, "gemv_fp4_grouped")
        return outs


def gemm_fp4(A_in: torch.Tensor, A: torch.Tensor, absmax: torch.Tensor,
             codebook: Optional[torch.Tensor], M: int, N: int, blocksize: int,
             bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Extension: dequant-fused tcgen05 GEMM, y = A_in @ W[M, N]^T (+ bias).  bf16 / fp16 only."""
    _check_in(A_in, "A_in")
    _check_in(A, "A", torch.uint8)
    _check_in(absmax, "absmax", torch.float32)
    if A_in.dtype not in (torch.float16, torch.bfloat16):
        raise RuntimeError("gemm_fp4 supports float16 and bfloat16 inputs")
    if A_in.shape[-1] != N:
        raise RuntimeError(f"A_in last dim {A_in.shape[-1]} != in_features {N}")
    rows = A_in.numel() // N if N else 0
    out = torch.empty(A_in.shape[:-1] + (M,), dtype=A_in.dtype, device=A_in.device)
    flags = 0
    if codebook is not None:
        _check_in(codebook, "codebook", torch.float32)
        if code_is_bnb_fp4(codebook):
            flags |= _lib.FLAG_CODE_IS_BNB_FP4
    if bias is not None:
        _check_in(bias, "bias", A_in.dtype)
    with _on_device(A_in) as st:
        check(lib.fp4_b200_gemm(
            A_in.data_ptr(), A.data_ptr(), absmax.data_ptr(),
            None if codebook is None else codebook.data_ptr(),
            None if bias is None else bias.data_ptr(), out.data_ptr(), rows, M, N, blocksize,
            _CODE_OF[A_in.dtype], flags, None, 0, st), "gemm_fp4")
    return out


def gemm_fp4_supported(rows: int, M: int, N: int, blocksize: int, dtype: torch.dtype) -> bool:
    return (dtype in (torch.float16, torch.bfloat16) and N % 64 == 0 and M % 8 == 0
            and blocksize % 64 == 0 and rows > 0 and GEMM_AVAILABLE)


GEMM_AVAILABLE = True  # the tcgen05 kernel is part of libfp4_b200.so


def _a_in_dtype(A_in: torch.Tensor) -> ScalarType:
    try:
        return ScalarType(_CODE_OF[A_in.dtype])
    except KeyError:
        raise RuntimeError(f"unsupported input dtype {A_in.dtype}") from None


def qlinear(A_in: torch.Tensor, A: torch.Tensor, absmax: torch.Tensor, M: int, N: int,
            blocksize: int) -> torch.Tensor:
    """reference csrc/torch_fp4.cpp:64-72: dequant (tree) to A_in's dtype, then linear."""
    _check_cuda(A_in, "A_in")
    w = dequantize_fp4(A, absmax, blocksize, M, N, _a_in_dtype(A_in))
    return torch.nn.functional.linear(A_in, w)


def qlinear_bias(A_in, A, absmax, M: int, N: int, blocksize: int, bias) -> torch.Tensor:
    """reference csrc/torch_fp4.cpp:74-82."""
    _check_cuda(A_in, "A_in")
    w = dequantize_fp4(A, absmax, blocksize, M, N, _a_in_dtype(A_in))
    return torch.nn.functional.linear(A_in, w, bias)


def qlinear_codebook(A_in, A, absmax, codebook, M: int, N: int, blocksize: int) -> torch.Tensor:
    """reference csrc/torch_fp4.cpp:84-92, with the whole weight dequantised."""
    _check_cuda(A_in, "A_in")
    w = dequantize_fp4_codebook(A, absmax, codebook, M, N, blocksize, M * N, _a_in_dtype(A_in))
    return torch.nn.functional.linear(A_in, w)


def qlinear_codebook_bias(A_in, A, absmax, codebook, M: int, N: int, blocksize: int,
                          bias) -> torch.Tensor:
    """reference csrc/torch_fp4.cpp:94-103, with the whole weight dequantised."""
    _check_cuda(A_in, "A_in")
    w = dequantize_fp4_codebook(A, absmax, codebook, M, N, blocksize, M * N, _a_in_dtype(A_in))
    return torch.nn.functional.linear(A_in, w, bias)


def quantize_fp4(w: torch.Tensor, blocksize: int = 64):
    """Extension: CUDA FP4 quantiser with the bitsandbytes thresholds.  Returns
    (packed uint8 [ceil(n/2), 1], absmax fp32 [ceil(n/blocksize)])."""
    _check_in(w, "w")
    if w.dtype not in _CODE_OF:
        raise RuntimeError(f"unsupported dtype {w.dtype}")
    n = w.numel()
    packed = torch.empty(((n + 1) // 2, 1), dtype=torch.uint8, device=w.device)
    absmax = torch.empty(((n + blocksize - 1) // blocksize,), dtype=torch.float32, device=w.device)
    with _on_device(w) as st:
        check(lib.fp4_b200_quantize(w.data_ptr(), _CODE_OF[w.dtype], n, blocksize,
                                    packed.data_ptr(), absmax.data_ptr(), st), "quantize_fp4")
    return packed, absmax
