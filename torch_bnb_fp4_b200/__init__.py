"""B200-native ``torch_bnb_fp4``: the reference's Python module surface (reference
torch_bnb_fp4/__init__.py:20-922) over hand-written sm_100a kernels.

Same public names, argument meaning and error behaviour as the reference for the quantised
``nn.Linear`` forward path; the kernels underneath are new.  What differs on purpose:

* ``QuantData.forward`` sends batch 1..8 inputs of any leading shape to the fused dequant-GEMV with
  the bias added in the kernel epilogue (the reference: batch 1 only, bias as a second op,
  :592-613), large inputs to the dequant-fused tcgen05 GEMM, everything else to dequant + cuBLAS;
* the compute dtype follows each call's input (the reference latches the first call's dtype, :590-591);
* nested (double-quantised) absmax is supported (the reference is not: README.md:223-224);
* every kernel runs on the current stream and is CUDA-graph capturable.

There is no CPU or PyTorch fallback for the kernels: importing this package without the built
``libfp4_b200.so`` raises.
"""
from __future__ import annotations

import logging
from enum import Enum
from math import prod
from typing import List, Optional, Tuple, TypeVar, Union

import torch
from torch import nn

from . import ext as _ext
from ._lib import FLAG_FORCE_GENERIC  # noqa: F401
from . import bnb_compat
from .bnb_compat import BF, HAVE_BNB, Linear4bit, LinearFP4, Params4bit  # noqa: F401
from .ext import (dequantize_fp4 as dequantize_fp4_, dequantize_fp4_codebook as dequantize_fp4_codebook_,
                  gemv_fp4 as gemv_fp4_, qlinear as qlinear_, qlinear_bias as qlinear_bias_,
                  qlinear_codebook as qlinear_codebook_, qlinear_codebook_bias as qlinear_codebook_bias_)
from .ext import ScalarType as ScalarType_

T_Model = TypeVar("T_Model", bound=nn.Module)

GEMV_MAX_BATCH = 8      # rows of x handled by the fused dequant-GEMV
GEMV_TWICE_MAX_WEIGHTS = 32 * 1024 * 1024  # 9..16 rows: layers up to this many weights take two 8-row GEMVs
GEMM_MIN_ROWS = 9       # rows of x from which the dequant-fused tcgen05 GEMM is used
GEMM_MAX_ROWS = 512     # ... and up to which it can beat new-dequant + cuBLAS on B200 (profiles/r01_gemm_sweep_*.log)
GEMM_SMALL_ROWS = 128   # up to here the fused GEMM wins (or ties) on every layer shape with K <= GEMM_LONG_K
GEMM_MID_MIN_WEIGHTS = 48 * 1024 * 1024  # 129..512 rows: only layers with at least this many weights
GEMM_LONG_K = 8192      # beyond this K a weight tile is a long serial k loop: the fused GEMM needs a full wave of them
GEMM_LONG_K_MIN_OUT = 148 * 128


def _fused_gemm_wins(rows: int, out_features: int, in_features: int) -> bool:
    """Which of the two bit-identical prefill paths is faster on a B200 (profiles/r02_gemm_sweep_shapes.log,
    r02_gemm_sweep_shapes_small_m.log, the gemm_sweep key of the bench line): the fused kernel runs one CTA per 128
    weight rows, so a layer with few row tiles (4096x4096: 32) leaves most SMs idle and a long K makes each tile a
    long serial loop (4096x14336: 0.6x at every M), while dequant + cuBLAS pays a full dequant pass that only large
    layers amortise.  Measured speed-up of the fused kernel: 1.0-1.8x for M <= 128 with K <= 8192 on every shape;
    for 256-512 rows 1.0-1.6x on 8192x8192 / 14336x4096 / 28672x8192 and 0.9x on 4096x4096 / 1024x4096; <= 1.03x
    beyond 512 rows everywhere."""
    if rows > GEMM_MAX_ROWS:
        return False
    if in_features > GEMM_LONG_K and out_features < GEMM_LONG_K_MIN_OUT:
        return False
    if rows <= GEMM_SMALL_ROWS:
        return True
    return out_features * in_features >= GEMM_MID_MIN_WEIGHTS


class ScalarType(Enum):
    """torch dtype <-> extension enum (reference torch_bnb_fp4/__init__.py:22-84)."""

    bfloat16 = ScalarType_.bfloat16
    float16 = ScalarType_.float16
    float32 = ScalarType_.float32

    @classmethod
    def from_torch_dtype(cls, dtype: torch.dtype) -> "ScalarType":
        if dtype == torch.bfloat16:
            return cls.bfloat16
        if dtype == torch.float16:
            return cls.float16
        if dtype == torch.float32:
            return cls.float32
        raise ValueError(f"Unsupported dtype {dtype}")

    @classmethod
    def from_str(cls, dtype: str) -> "ScalarType":
        try:
            return {"bfloat16": cls.bfloat16, "float16": cls.float16, "float32": cls.float32}[dtype]
        except KeyError:
            raise ValueError(f"Unsupported dtype {dtype}") from None

    @property
    def torch_dtype(self) -> torch.dtype:
        # the reference's property names non-existent members (:77-82); this one works
        return {ScalarType.bfloat16: torch.bfloat16, ScalarType.float16: torch.float16,
                ScalarType.float32: torch.float32}[self]


# ---- thin op shims (reference :87-337): same names, same argument order --------------------------
@torch.no_grad()
def dequantize_fp4(qweight, absmax, blocksize: int, M: int, N: int, dtype=torch.float16):
    return dequantize_fp4_(qweight, absmax, blocksize, M, N, ScalarType.from_torch_dtype(dtype).value)


@torch.no_grad()
def dequantize_fp4_qtype(qweight, absmax, blocksize: int, M: int, N: int,
                         dtype=ScalarType.bfloat16.value):
    return dequantize_fp4_(qweight, absmax, blocksize, M, N, dtype)


@torch.no_grad()
def dequantize_fp4_codebook_invoke_qtype(qweight, absmax, code, blocksize: int, M: int, N: int,
                                         numel: int, qtype):
    return dequantize_fp4_codebook_(qweight, absmax, code, M, N, blocksize, numel, qtype)


@torch.no_grad()
def dequantize_fp4_codebook_invoke(qweight, absmax, code, blocksize: int, M: int, N: int,
                                   numel: int, qtype: torch.dtype):
    return dequantize_fp4_codebook_(qweight, absmax, code, M, N, blocksize, numel,
                                    ScalarType.from_torch_dtype(qtype).value)


@torch.no_grad()
def gemm_4bit_inference(A, B, absmax, code, blocksize: int, dtype=torch.float16, Bshape=None):
    return gemv_fp4_(A, B, absmax, code, blocksize, ScalarType.from_torch_dtype(dtype).value, Bshape)


@torch.no_grad()
def gemm_4bit_inference_qtype(A, B, absmax, code, blocksize: int,
                              dtype=ScalarType.bfloat16.value, Bshape: List[int] = None):
    return gemv_fp4_(A, B, absmax, code, blocksize, dtype, Bshape)


class QuantData:
    """Packed weight + quantisation state of one linear layer, and the forward dispatcher
    (reference torch_bnb_fp4/__init__.py:340-618)."""

    def __init__(self, A: torch.Tensor, state, shape: Tuple[int, int], original_lin,
                 bias: Optional[torch.Tensor] = None, use_codebook_dequant: Optional[bool] = True,
                 allow_reduced_precision_linear: Optional[bool] = False,
                 materialize_nested_absmax: bool = False):
        self.use_codebook_dequant = use_codebook_dequant
        self.A = A
        self.blocksize = state.blocksize
        self.M = shape[0]  # out_features
        self.N = shape[1]  # in_features
        self.code = state.code.float().contiguous()
        self.quant_state = state
        self.original_lin = original_lin
        self.numel = prod(shape)
        self.nested = None
        nblocks = (self.numel + self.blocksize - 1) // self.blocksize
        if getattr(state, "nested", False):
            # double-quantised absmax (reference: unsupported, README.md:223-224)
            s2 = state.state2
            offset = float(state.offset) if not torch.is_tensor(state.offset) else float(state.offset.item())
            self.nested = _ext.make_nested(state.absmax.contiguous(), s2.code.float().contiguous(),
                                           s2.absmax.float().contiguous(), offset, s2.blocksize)
            self.absmax = None
            if materialize_nested_absmax:
                self.absmax = _ext.absmax_denest(self.nested, nblocks, A.device)
                self.nested = None
        else:
            self.absmax = state.absmax.float().contiguous()
        self.bias = original_lin.bias if hasattr(original_lin, "bias") else bias
        self._bias_by_dtype = {}
        self._fast = None
        self.o_type = None
        self.qtype = None
        self.compute_dtype_set = False
        self.allow_reduced_precision_linear = allow_reduced_precision_linear
        # the reference's "reduced precision" variants (:391-396) are the same dequant + linear done
        # inside the extension; kept selectable, computed correctly
        if allow_reduced_precision_linear and self.nested is None:
            self.qlinear = (self._qlinear_low_precision_codebook if use_codebook_dequant
                            else self._qlinear_low_precision_normal)
        else:
            self.qlinear = self._dequant_linear
        self.dequantize = self._dequantize_codebook if use_codebook_dequant else self._dequantize_normal
        self._code_is_std = _ext.code_is_bnb_fp4(self.code)  # one device->host read, at load time
        self._Bshape = (self.M, self.N)

    # -- dtype handling ---------------------------------------------------------------------------
    def set_compute_type(self, x: torch.Tensor) -> None:
        self.o_type = x.dtype
        self.qtype = ScalarType.from_torch_dtype(x.dtype).value
        if self.bias is not None:
            b = self._bias_by_dtype.get(x.dtype)
            if b is None:
                b = self.bias.detach().to(dtype=x.dtype).contiguous()
                self._bias_by_dtype[x.dtype] = b
            self._bias_t = b
        else:
            self._bias_t = None
        self.compute_dtype_set = True
        # pre-validated launcher for the decode GEMV of this dtype (None: take the checked path)
        self._fast = None
        if self.nested is None and self.absmax is not None and self.blocksize % 32 == 0 and self.N % 32 == 0:
            try:
                self._fast = _ext.GemvLauncher(self.A, self.absmax, self.code, self.blocksize, self.qtype,
                                               self._Bshape, self._bias_t)
                self._fast_idx = self.A.device.index
            except Exception:  # noqa: BLE001 - anything unusual: the checked path reports it properly
                self._fast = None

    # -- dequant ----------------------------------------------------------------------------------
    def _dequantize_codebook(self) -> torch.Tensor:
        if self.nested is not None:
            return _ext.dequantize_fp4_nested(self.A, self.nested, self.code, self.M, self.N,
                                              self.blocksize, self.qtype)
        return dequantize_fp4_codebook_invoke_qtype(self.A, self.absmax, self.code, self.blocksize,
                                                    self.M, self.N, self.numel, self.qtype)

    def _dequantize_normal(self) -> torch.Tensor:
        if self.nested is not None:
            return _ext.dequantize_fp4_nested(self.A, self.nested, None, self.M, self.N,
                                              self.blocksize, self.qtype)
        return dequantize_fp4_qtype(self.A, self.absmax, self.blocksize, self.M, self.N, self.qtype)

    def _dequant_linear(self, A: torch.Tensor) -> torch.Tensor:
        return torch.nn.functional.linear(A, self.dequantize(), self._bias_t)

    def _qlinear_low_precision_normal(self, A: torch.Tensor) -> torch.Tensor:
        if self._bias_t is None:
            return qlinear_(A, self.A, self.absmax, self.M, self.N, self.blocksize)
        return qlinear_bias_(A, self.A, self.absmax, self.M, self.N, self.blocksize, self._bias_t)

    def _qlinear_low_precision_codebook(self, A: torch.Tensor) -> torch.Tensor:
        if self._bias_t is None:
            return qlinear_codebook_(A, self.A, self.absmax, self.code, self.M, self.N, self.blocksize)
        return qlinear_codebook_bias_(A, self.A, self.absmax, self.code, self.M, self.N,
                                      self.blocksize, self._bias_t)

    # -- fused paths ------------------------------------------------------------------------------
    def _qgemv(self, A: torch.Tensor) -> torch.Tensor:
        """Fused dequant-GEMV, bias in the epilogue.  A: [..., K] with <= 8 rows."""
        return _ext.gemv_fp4_bias(A, self.A, self.absmax, self.code, self.blocksize, self.qtype,
                                  self._Bshape, self._bias_t, self.nested)

    def _qgemm(self, A: torch.Tensor) -> torch.Tensor:
        return _ext.gemm_fp4(A, self.A, self.absmax, self.code, self.M, self.N, self.blocksize,
                             self._bias_t)

    def forward(self, A: torch.Tensor) -> torch.Tensor:
        # decode fast path first (1-2 rows always fit the GEMV; everything else was validated once in
        # set_compute_type): as few interpreter steps as possible between the module call and the launch
        f = self._fast
        if f is not None and A.dtype is self.o_type and A.shape[-1] == self.N:
            n_el = A.numel()
            if ((n_el == self.N or n_el == 2 * self.N) and A.is_contiguous()
                    and A.device.index == self._fast_idx == torch.cuda.current_device()):
                return f(A, 1 if n_el == self.N else 2)
        k = A.shape[-1]
        n_el = A.numel()
        if n_el == 0:  # same shapes as the reference's empty-input branch (:580-589)
            B_shape = self.quant_state.shape
            tail = B_shape[1:] if k == B_shape[0] else B_shape[:1]
            return torch.empty(A.shape[:-1] + tail, dtype=A.dtype, device=A.device)
        if A.dtype != self.o_type:
            self.set_compute_type(A)
        rows = n_el // k
        gemm_ok = (self.nested is None and self._code_is_std
                   and _ext.gemm_fp4_supported(rows, self.M, self.N, self.blocksize, A.dtype))
        if rows <= GEMV_MAX_BATCH and k % 32 == 0 and self.blocksize % 32 == 0:
            # the streaming GEMV keeps x (as integer terms) in shared memory; where a batch does not fit (8 rows x
            # K = 14336, fp32 5..8 rows x K = 8192) the C-ABI runs it as two launches of half the rows each - still
            # 2-3x faster than the tensor-core GEMM with a 16-token tile (88 us on 4096x14336)
            if not A.is_contiguous():
                A = A.contiguous()
            return self._qgemv(A)
        if (GEMV_MAX_BATCH < rows <= 2 * GEMV_MAX_BATCH and self.numel <= GEMV_TWICE_MAX_WEIGHTS
                and k % 32 == 0 and self.blocksize % 32 == 0 and self._code_is_std):
            # 9..16 rows on a small layer: the GEMM's few weight tiles leave most SMs idle and its one dequantiser
            # warp per sub-partition sets the pace (16 us on 2048x2048); two 8-row GEMVs are faster there
            A2 = A.reshape(rows, k)
            if not A2.is_contiguous():
                A2 = A2.contiguous()
            f = self._fast  # the pre-validated launcher where there is one (the checked path costs ~20 us per call)
            if (f is not None and A2.dtype is self.o_type and A2.device.index == self._fast_idx == torch.cuda.current_device()):
                y = torch.cat([f(A2[:GEMV_MAX_BATCH], GEMV_MAX_BATCH), f(A2[GEMV_MAX_BATCH:], rows - GEMV_MAX_BATCH)], dim=0)
            else:
                y = torch.cat([self._qgemv(A2[:GEMV_MAX_BATCH]), self._qgemv(A2[GEMV_MAX_BATCH:])], dim=0)
            return y.view(A.shape[:-1] + (self.M,))
        if gemm_ok and _fused_gemm_wins(rows, self.M, self.N):
            if not A.is_contiguous():
                A = A.contiguous()
            return self._qgemm(A)
        return self.qlinear(A)


class TorchFP4Linear(nn.Module):
    """Wrapper for a quantised bitsandbytes LinearFP4 / Linear4bit (reference :621-714)."""

    def __init__(self, lin, use_codebook_dequant: bool = True, name: str = "",
                 materialize_nested_absmax: bool = True):
        # materialize_nested_absmax (extension): a double-quantised absmax is decoded ONCE at load into
        # fp32 (bit-exact) so decode takes the streaming GEMV; False keeps it nested and decodes
        # it inside the kernels (0.52 instead of 0.56 bytes per weight, slower kernels)
        super().__init__()
        self.lin = [lin]
        self.in_features = lin.in_features
        self.out_features = lin.out_features
        self.use_codebook_dequant = use_codebook_dequant
        self.name = name
        self._materialize_nested = materialize_nested_absmax
        w = lin.weight
        if not (isinstance(w, Params4bit) or hasattr(w, "quant_state")):
            raise ValueError("Linear is not a bnb linear and is not quantized, and I have no idea "
                             "what to do with that rn.")
        if (w.quant_state is None or w.device.type != "cuda" or w.data.dtype != torch.uint8):
            raise ValueError("Linear weights are not quantized, and I have no idea what to do with "
                             f"that rn. Weights are {w.data.dtype}")
        qtype = getattr(w.quant_state, "quant_type", "fp4")
        if qtype != "fp4":
            raise ValueError(f"only quant_type='fp4' is supported, got {qtype!r}")
        self.quant_data = QuantData(w.data, w.quant_state, w.quant_state.shape, bias=lin.bias,
                                    original_lin=lin, use_codebook_dequant=use_codebook_dequant,
                                    materialize_nested_absmax=materialize_nested_absmax)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.quant_data.forward(x)

    def __repr__(self) -> str:
        lin = self.lin[0]
        dt = f", dtype={self.quant_data.o_type}" if hasattr(self, "quant_data") else ""
        return (f"TorchFP4Linear(in_features={lin.in_features}, out_features={lin.out_features}, "
                f"bias={lin.bias is not None}{dt})")

    # -- on-disk / wire format (the reference keeps the layer in a python list and has no
    #    state_dict at all): the keys bitsandbytes' Linear4bit writes, so checkpoints are interchangeable
    def quantized_state_dict(self, prefix: str = "") -> dict:
        lin = self.lin[0]
        out = {prefix + "weight": lin.weight.data}
        for k, v in bnb_compat.quant_state_as_dict(lin.weight.quant_state, packed=True).items():
            out[prefix + "weight." + k] = v
        if getattr(lin, "bias", None) is not None:
            out[prefix + "bias"] = lin.bias.data
        return out

    @classmethod
    def from_quantized_state_dict(cls, sd: dict, prefix: str = "", device="cuda", tp_rank: int = 0,
                                  tp_world: int = 1, tp_mode: Optional[str] = None, **kw) -> "TorchFP4Linear":
        """Rebuild a layer from the bitsandbytes 4-bit keys (`weight`, `weight.absmax`, `weight.quant_map`,
        `weight.quant_state.bitsandbytes__fp4`, optional `weight.nested_*`, `bias`).

        With ``tp_world > 1`` only this rank's shard is built: ``tp_mode="column"`` keeps rows
        [rank*N/tp, (rank+1)*N/tp) (q/k/v/gate/up), ``"row"`` keeps columns [rank*K/tp, ...) of every row (o/down;
        the bias stays with rank 0, which adds it once before the reduction).  The cut is made on the HOST copy
        of the checkpoint tensors, so the full weight never reaches the GPU; a nested absmax is materialised
        first because its 256-block grouping does not line up with shard boundaries."""
        w = sd[prefix + "weight"]
        comp = {k[len(prefix) + len("weight."):]: v for k, v in sd.items() if k.startswith(prefix + "weight.")}
        bias = sd.get(prefix + "bias")
        if tp_world > 1:
            from .parallel import shard_column, shard_row
            if tp_mode not in ("column", "row"):
                raise ValueError("tp_mode must be 'column' or 'row' when tp_world > 1")
            qs_full = bnb_compat.quant_state_from_dict(comp, device=None)
            N, K, bs = int(qs_full.shape[0]), int(qs_full.shape[1]), int(qs_full.blocksize)
            absmax = qs_full.absmax
            if qs_full.nested:  # decode on the device (bit-exact kernel), slice on the host
                nd = _ext.make_nested(absmax.to(device).contiguous(), qs_full.state2.code.float().to(device).contiguous(),
                                      qs_full.state2.absmax.float().to(device).contiguous(),
                                      float(qs_full.offset), qs_full.state2.blocksize)
                absmax = _ext.absmax_denest(nd, (N * K + bs - 1) // bs, torch.device(device)).cpu()
            absmax = absmax.float()
            if tp_mode == "column":
                p, a, n = shard_column(w, absmax, N, K, tp_rank, tp_world, bs)
                shape = (n, K)
                if bias is not None:
                    bias = bias[tp_rank * n:(tp_rank + 1) * n]
            else:
                p, a, k = shard_row(w, absmax, N, K, tp_rank, tp_world, bs)
                shape = (N, k)
                if tp_rank != 0:
                    bias = None
            qs = bnb_compat.QuantState(absmax=a.to(device), shape=torch.Size(shape), code=qs_full.code.to(device),
                                       blocksize=bs, quant_type="fp4", dtype=qs_full.dtype)
            w = p
        else:
            qs = bnb_compat.quant_state_from_dict(comp, device=device)
        out_f, in_f = int(qs.shape[0]), int(qs.shape[1])
        lin = bnb_compat.LinearFP4(in_f, out_f, bias=bias is not None, compress_statistics=bool(qs.nested))
        lin.weight = Params4bit(w.to(device).contiguous().view(-1, 1), requires_grad=False, quant_state=qs,
                                blocksize=qs.blocksize, compress_statistics=bool(qs.nested), quant_type="fp4")
        if bias is not None:
            lin.bias = nn.Parameter(bias.detach().clone().to(device), requires_grad=False)
        return cls(lin, **kw)

    # nn.Module.state_dict() / load_state_dict(): the same keys, written and read the way bitsandbytes' Linear4bit
    # does (_save_to_state_dict adds the quant-state components next to `weight`), so a converted model saves and
    # reloads with torch.save(model.state_dict()) like any other module.
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for k, v in self.quantized_state_dict(prefix).items():
            destination[k] = v if keep_vars else v.detach()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        if prefix + "weight" not in state_dict:
            missing_keys.append(prefix + "weight")
            return
        try:
            new = TorchFP4Linear.from_quantized_state_dict(
                state_dict, prefix, device=self.quant_data.A.device, use_codebook_dequant=self.use_codebook_dequant,
                name=self.name, materialize_nested_absmax=self._materialize_nested)
        except Exception as e:  # noqa: BLE001
            error_msgs.append(f"While loading {prefix}weight: {e}")
            return
        if (new.in_features, new.out_features) != (self.in_features, self.out_features):
            error_msgs.append(f"size mismatch for {prefix}weight: checkpoint holds "
                              f"{new.out_features}x{new.in_features}, the module is "
                              f"{self.out_features}x{self.in_features}")
            return
        self.lin, self.quant_data = new.lin, new.quant_data

    @classmethod
    def from_linear(cls, linear, use_codebook_dequant: bool = False, name: str = "") -> "TorchFP4Linear":
        return cls(linear, use_codebook_dequant=use_codebook_dequant, name=name)


class TorchFP4LinearGroup(nn.Module):
    """Extension: several TorchFP4Linear layers that consume the SAME input - the
    q/k/v or gate/up projections of a decoder layer - evaluated with one fused dequant-GEMV launch for
    decode-sized inputs.  forward(x) returns the tuple of outputs, each equal to calling the layer on its own
    (up to fp32 summation order); anything the grouped kernel does not cover falls back to exactly that."""

    def __init__(self, layers):
        super().__init__()
        self.layers = nn.ModuleList(layers)
        qds = [m.quant_data for m in self.layers]
        # (a member may keep its absmax double-quantised: the kernel decodes it)
        self._groupable = (len(qds) <= 4 and all(q._code_is_std and q.blocksize == 64 for q in qds)
                           and len({q.N for q in qds}) == 1)
        self._launchers = {}

    def forward(self, x: torch.Tensor):
        qds = [m.quant_data for m in self.layers]
        k = x.shape[-1]
        rows = x.numel() // k if k else 0
        if self._groupable and 0 < rows <= GEMV_MAX_BATCH and k == qds[0].N:
            for q in qds:
                if x.dtype != q.o_type:
                    q.set_compute_type(x)
            # pre-validated launcher per input dtype (pointer arrays built once)
            if x.is_cuda and x.is_contiguous() and x.device.index == torch.cuda.current_device():
                la = self._launchers.get(x.dtype)
                if la is None:
                    try:
                        la = _ext.GroupLauncher([q.A for q in qds], [q.absmax for q in qds], 64, qds[0].qtype,
                                                [q._Bshape for q in qds], [q._bias_t for q in qds],
                                                [q.nested for q in qds])
                    except Exception:  # noqa: BLE001 - the checked path below reports it properly
                        la = False
                    self._launchers[x.dtype] = la
                if la and la.idx == x.device.index:
                    outs = la(x, rows)
                    if outs is not None:
                        return tuple(outs)
            if all(q.nested is None for q in qds):
                xc = x if x.is_contiguous() else x.contiguous()
                outs = _ext.gemv_fp4_grouped(xc, [q.A for q in qds], [q.absmax for q in qds], 64, qds[0].qtype,
                                             [q._Bshape for q in qds], [q._bias_t for q in qds])
                if outs is not None:
                    return tuple(outs)
        return tuple(m(x) for m in self.layers)


def _unwrap(m):
    return m.layer if isinstance(m, _GroupMember) else m


def _decode_ready(qds, x) -> bool:
    """the fused epilogues cover what the streaming GEMV covers: bitsandbytes table, blocksize 64, fp32 absmax,
    1..8 rows of a contiguous CUDA input"""
    k = x.shape[-1]
    rows = x.numel() // k if k else 0
    return (0 < rows <= GEMV_MAX_BATCH and x.is_cuda and all(
        q.nested is None and q.absmax is not None and q._code_is_std and q.blocksize == 64 and q.N == k for q in qds))


def linear_add(layer, x: torch.Tensor, residual: torch.Tensor) -> torch.Tensor:
    """``layer(x) + residual`` with the addition in the GEMV epilogue for decode-sized inputs (the residual stream
    around an o / down projection: one launch instead of two, the sum rounded once)."""
    layer = _unwrap(layer)
    qd = layer.quant_data
    if _decode_ready([qd], x) and residual.dtype == x.dtype:
        if x.dtype != qd.o_type:
            qd.set_compute_type(x)
        xc = x if x.is_contiguous() else x.contiguous()
        outs = _ext.gemv_fp4_fused(xc, [qd.A], [qd.absmax], 64, qd.qtype, [qd._Bshape], [qd._bias_t],
                                   residuals=[residual.contiguous()])
        if outs is not None:
            return outs[0]
    return layer(x) + residual


class TorchFP4GatedMLP(nn.Module):
    """Extension: ``down(act(gate(x)) * up(x)) [+ residual]`` of a Llama / Mistral MLP block.
    For decode-sized inputs the gate and up projections run as ONE launch whose epilogue applies the activation
    and the product (the intermediate [rows, inter] tensors of the unfused form are never written), and the
    residual is added in the down projection's epilogue: two launches instead of three GEMVs plus two or three
    elementwise kernels.  Anything else (prefill) composes the layers' own paths."""

    def __init__(self, gate, up, down, act: str = "silu"):
        super().__init__()
        if act not in ("silu", "gelu_tanh"):
            raise ValueError("act must be 'silu' or 'gelu_tanh'")
        self.gate_proj, self.up_proj, self.down_proj, self.act = _unwrap(gate), _unwrap(up), _unwrap(down), act

    def _act(self, v):
        return torch.nn.functional.silu(v) if self.act == "silu" else torch.nn.functional.gelu(v, approximate="tanh")

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        g, u = self.gate_proj.quant_data, self.up_proj.quant_data
        if _decode_ready([g, u], x) and g.M == u.M:
            for q in (g, u):
                if x.dtype != q.o_type:
                    q.set_compute_type(x)
            xc = x if x.is_contiguous() else x.contiguous()
            outs = _ext.gemv_fp4_fused(xc, [g.A, u.A], [g.absmax, u.absmax], 64, g.qtype, [g._Bshape, u._Bshape],
                                       [g._bias_t, u._bias_t], gate_act=self.act)
            if outs is not None:
                h = outs[0]
                return self.down_proj(h) if residual is None else linear_add(self.down_proj, h, residual)
        h = self._act(self.gate_proj(x)) * self.up_proj(x)
        y = self.down_proj(h)
        return y if residual is None else y + residual


def _which_activation(fn) -> Optional[str]:
    """'silu' / 'gelu_tanh' if `fn` computes that function (probed numerically, so nn.SiLU, F.silu and the
    activation classes of transformers are all recognised), else None."""
    if fn is None or not callable(fn):
        return None
    t = torch.linspace(-6.0, 6.0, 97)
    try:
        with torch.no_grad():
            y = fn(t.clone())
    except Exception:  # noqa: BLE001
        return None
    if not torch.is_tensor(y) or y.shape != t.shape:
        return None
    if torch.allclose(y, torch.nn.functional.silu(t), atol=1e-6, rtol=1e-5):
        return "silu"
    if torch.allclose(y, torch.nn.functional.gelu(t, approximate="tanh"), atol=1e-6, rtol=1e-5):
        return "gelu_tanh"
    return None


def fuse_gated_mlps(model: nn.Module) -> int:
    """Opt-in: give every sub-module that looks like an HF gated MLP - TorchFP4Linear children named gate_proj /
    up_proj / down_proj and an ``act_fn`` that computes SiLU or tanh-GELU (probed numerically), forward =
    down(act(gate(x)) * up(x)) - a forward
    that runs through TorchFP4GatedMLP.  Returns the number of blocks fused."""
    import types
    made = 0
    for mod in model.modules():
        subs = [getattr(mod, n, None) for n in ("gate_proj", "up_proj", "down_proj")]
        if not all(isinstance(_unwrap(m), TorchFP4Linear) for m in subs if m is not None) or None in subs:
            continue
        act = _which_activation(getattr(mod, "act_fn", None))
        if act is None:
            continue
        fused = TorchFP4GatedMLP(*subs, act=act)
        mod.__dict__["_fp4_fused_mlp"] = fused  # not registered: the block keeps owning its layers
        mod.forward = types.MethodType(lambda self, x: self.__dict__["_fp4_fused_mlp"](x), mod)
        made += 1
    return made


class _GroupMember(nn.Module):
    """One projection of a TorchFP4LinearGroup that is called like a plain nn.Linear: the FIRST member called with
    a new input runs the whole group in one launch and parks the siblings' outputs; the siblings return them when
    they are called with the same input tensor object (unmodified since) and fall back to their own launch
    otherwise.  This is what lets an unmodified model (``self.q_proj(x); self.k_proj(x); self.v_proj(x)``) use
    the grouped kernel."""

    def __init__(self, group: "TorchFP4LinearGroup", index: int):
        super().__init__()
        self._group = [group]  # hidden from module registration: the group owns the layers
        self.index = index
        self.layer = group.layers[index]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        g = self._group[0]
        # hit: the SAME tensor object, unmodified since the grouped launch (the usual pattern: one hidden_states
        # passed to q_proj, k_proj, v_proj in turn); anything else simply launches on its own
        if g._cache_x is x and g._cache_ver == x._version:
            out = g._cache[self.index]
            if out is not None:
                g._cache[self.index] = None
                if all(o is None for o in g._cache):  # last sibling served: drop the references
                    g._cache_x = None
                return out
        k = x.shape[-1]
        rows = x.numel() // k if k else 0
        if not (0 < rows <= GEMV_MAX_BATCH):
            return self.layer(x)
        outs = list(g(x))
        g._cache_x, g._cache_ver = x, x._version  # holding x keeps its storage from being reused meanwhile
        out, outs[self.index] = outs[self.index], None
        g._cache = outs
        return out

    def __getattr__(self, name):  # in_features, weight, ... resolve on the wrapped layer
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("layer"), name)


_GROUPABLE_NAMES = (("q_proj", "k_proj", "v_proj"), ("gate_proj", "up_proj"))


def group_projections(model: nn.Module, names=_GROUPABLE_NAMES) -> int:
    """Extension: find sibling TorchFP4Linear layers that read the same input in decoder blocks (q/k/v,
    gate/up by their usual names) and make them share one grouped launch, without touching the model's
    forward code.  Returns the number of groups created."""
    made = 0
    for parent in model.modules():
        for group_names in names:
            subs = [getattr(parent, n, None) for n in group_names]
            if not all(isinstance(m, TorchFP4Linear) for m in subs):
                continue
            if len({m.in_features for m in subs}) != 1:
                continue
            grp = TorchFP4LinearGroup(subs)
            grp._cache_x, grp._cache_ver, grp._cache = None, -1, [None] * len(subs)
            if not grp._groupable:
                continue
            for i, n in enumerate(group_names):
                setattr(parent, n, _GroupMember(grp, i))

            # parked sibling outputs never outlive the block's forward: a later call that happens to pass the
            # same tensor object again (a static buffer refilled through a raw pointer, which the version counter
            # does not see) starts from a fresh launch
            def _drop(_m, _i, _o, g=grp):
                g._cache_x, g._cache = None, [None] * len(g._cache)
            parent.register_forward_hook(_drop)
            made += 1
    return made


@torch.no_grad()
def swap_linear_with_bnb_linear(linear: nn.Linear, dtype=torch.float16):
    """nn.Linear -> (unquantised) LinearFP4 carrying the same weights (reference :717-747)."""
    mod = LinearFP4(linear.in_features, linear.out_features, bias=linear.bias is not None,
                    compute_dtype=dtype)
    mod.weight.data = linear.weight.data.clone().detach()
    if linear.bias is not None:
        mod.bias.data = linear.bias.data.clone().detach()
    mod.requires_grad_(False)
    return mod


def check_if_name_contained_in_list(name, names_list):
    return any(n in name for n in names_list)


def todevice_if_necessary(module, device):
    """Make sure a bnb linear is on `device` AND quantised (reference :759-778)."""
    if module.weight.data.dtype != torch.uint8 and isinstance(module, (Linear4bit, LinearFP4)):
        module.weight = module.weight.to(device)
        if not (module.weight.data.device == torch.device(device)
                and module.weight.data.dtype == torch.uint8):
            logging.debug("layer reached the device unquantised; quantising it directly")
            qweight, qstate = BF.quantize_fp4(module.weight.data)
            module.weight.data = qweight
            module.weight.quant_state = qstate
    return module


def _to_fp4(module, device, as_dtype, use_codebook_dequant, name):
    if not isinstance(module, (LinearFP4, Linear4bit)):
        module = swap_linear_with_bnb_linear(module, dtype=as_dtype)
    module = module.to(device)
    if getattr(module.weight, "quant_state", None) is None:
        module = todevice_if_necessary(module, device)
    return TorchFP4Linear(lin=module, use_codebook_dequant=use_codebook_dequant, name=name)


def recursively_replace_with_fp4_linear(
    module: T_Model, as_dtype=torch.float16, use_codebook_dequant=True,
    device: torch.device = torch.device("cuda" if torch.cuda.is_available() else "cpu"),
    return_final_module: bool = True, only_replace_bnb_layers: bool = False,
    ignore_layer_names: List[str] = ["lm_head"], parent="", debug: bool = False,
    group_projections_: bool = True, _memo: Optional[dict] = None,
) -> Optional[T_Model]:
    """Swap every nn.Linear / LinearFP4 / Linear4bit below `module` for a TorchFP4Linear
    (reference :781-922; same keyword arguments and defaults).

    Two differences, both on purpose.  (1) The reference walks ``named_children()``, which yields a module that is
    registered under several names only ONCE: in ``nn.Sequential(*([act, lin] * 4))`` (its own sanity model,
    sanity_check.py:42-44) only the first alias is swapped and the other three keep calling the unquantised
    nn.Linear.  Here every alias gets the same replacement.  (2) ``group_projections_`` (default on, top-level call
    only): sibling q/k/v and gate/up projections of decoder blocks share one fused launch afterwards
    (``group_projections``); outputs are the same up to fp32 summation order."""
    dev_type = device.type if hasattr(device, "type") else str(device).split(":")[0]
    assert dev_type == "cuda", "Device type must be cuda!"
    prefix = parent + "." if parent != "" else ""
    swapped_plain = False
    memo = {} if _memo is None else _memo  # id(original module) -> its replacement: aliases stay aliases
    for name, child in list(module._modules.items()):
        if child is None:
            continue
        child_name = prefix + name
        if check_if_name_contained_in_list(name, ignore_layer_names):
            if debug:
                print(f"Ignoring name: {child_name}, as it is in the ignore list")
            continue
        if id(child) in memo:
            module._modules[name] = memo[id(child)]
            continue
        if isinstance(child, (LinearFP4, Linear4bit)):
            if debug:
                print(f"Replacing BNB layer {child_name} swapping with TorchFP4Linear.")
            memo[id(child)] = module._modules[name] = _to_fp4(child, device, as_dtype, use_codebook_dequant, child_name)
        elif isinstance(child, nn.Linear):
            if only_replace_bnb_layers:
                if debug:
                    print(f"Ignoring {child_name}, as only_replace_bnb_layers=True")
            else:
                if debug:
                    print(f"Replacing {child_name} with BNB linear, and then swapping with TorchFP4Linear.")
                memo[id(child)] = module._modules[name] = _to_fp4(child, device, as_dtype, use_codebook_dequant,
                                                                  child_name)
                swapped_plain = True
        elif isinstance(child, nn.Module):
            memo[id(child)] = child  # visited: a shared sub-tree is converted once
            recursively_replace_with_fp4_linear(
                child, as_dtype=as_dtype, use_codebook_dequant=use_codebook_dequant, device=device,
                return_final_module=False, only_replace_bnb_layers=only_replace_bnb_layers,
                ignore_layer_names=ignore_layer_names, parent=child_name, debug=debug,
                group_projections_=False, _memo=memo)
    if isinstance(module, (LinearFP4, Linear4bit)):
        module = _to_fp4(module, device, as_dtype, use_codebook_dequant, parent)
    elif isinstance(module, nn.Linear) and not only_replace_bnb_layers:
        module = _to_fp4(module, device, as_dtype, use_codebook_dequant, parent)
        swapped_plain = True
    if swapped_plain:
        torch.cuda.empty_cache()
    if group_projections_ and _memo is None and isinstance(module, nn.Module):
        group_projections(module)
    if return_final_module:
        return module
