"""bitsandbytes stand-ins.

The reference imports ``bitsandbytes`` for the objects on either side of the hot path
(``Params4bit`` / ``QuantState`` / ``LinearFP4`` / ``Linear4bit`` / ``functional.quantize_fp4``,
reference torch_bnb_fp4/__init__.py:7-9).  bitsandbytes is not installed in this image, so the
package duck-types those objects: when ``import bitsandbytes`` works its classes are used, otherwise
the minimal equivalents below (same attribute names, same packed layout) are.  Quantisation runs on
the GPU through ``fp4_b200_quantize`` (bitsandbytes thresholds); there is no CPU path.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

try:  # pragma: no cover - not available in this image
    import bitsandbytes as _bnb  # type: ignore
    from bitsandbytes import functional as BF  # type: ignore
    from bitsandbytes.nn.modules import Linear4bit, LinearFP4, Params4bit  # type: ignore
    from bitsandbytes.functional import QuantState  # type: ignore

    HAVE_BNB = True
except Exception:  # noqa: BLE001
    HAVE_BNB = False

from . import ext

FP4_CODE = torch.tensor(ext.BNB_FP4_CODE, dtype=torch.float32)


def create_dynamic_map(signed: bool = True, max_exponent_bits: int = 7, total_bits: int = 8) -> torch.Tensor:
    """bitsandbytes' 8-bit "dynamic" code used for the nested absmax (restated; the kernels only
    gather from whatever 256-entry table the quant_state carries)."""
    data = []
    non_sign_bits = total_bits - (1 if signed else 0)
    additional_items = 2 ** (non_sign_bits - max_exponent_bits) - 1
    i = 0
    for i in range(max_exponent_bits):
        fraction_items = int(2 ** (i + non_sign_bits - max_exponent_bits) + 1 if signed
                             else 2 ** (i + non_sign_bits - max_exponent_bits + 1) + 1)
        boundaries = torch.linspace(0.1, 1, fraction_items)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        data += ((10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
        if signed:
            data += (-(10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
    if additional_items > 0:
        boundaries = torch.linspace(0.1, 1, additional_items + 1)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        data += ((10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
        if signed:
            data += (-(10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
    data.append(0)
    data.append(1.0)
    data += [0] * (2 ** total_bits - len(data))
    data.sort()
    return torch.tensor(data, dtype=torch.float32)


def quantize_blockwise_8bit(v: torch.Tensor, code: torch.Tensor, blocksize: int = 256):
    """Nearest-entry 8-bit blockwise quantisation (bitsandbytes quantize_blockwise semantics);
    load-time only, plain torch ops on the tensor's device."""
    n = v.numel()
    nblk = (n + blocksize - 1) // blocksize
    pad = nblk * blocksize - n
    vp = torch.nn.functional.pad(v.float().flatten(), (0, pad)).view(nblk, blocksize)
    absmax2 = vp.abs().amax(dim=1)
    normed = vp / absmax2.clamp_min(1e-38).unsqueeze(1)
    code = code.to(v.device)
    mid = (code[1:] + code[:-1]) / 2
    q = torch.bucketize(normed.contiguous(), mid).to(torch.uint8).flatten()[:n].contiguous()
    return q, absmax2.contiguous()


if not HAVE_BNB:

    class QuantState:
        """Field-compatible with bitsandbytes.functional.QuantState (the fields the reference
        reads: torch_bnb_fp4/__init__.py:377-390)."""

        def __init__(self, absmax, shape=None, code=None, blocksize=64, quant_type="fp4",
                     dtype=torch.float16, offset=None, state2=None):
            self.absmax = absmax
            self.shape = torch.Size(shape) if shape is not None else None
            self.code = code
            self.blocksize = blocksize
            self.quant_type = quant_type
            self.dtype = dtype
            self.offset = offset
            self.state2 = state2
            self.nested = state2 is not None

        def to(self, device):
            self.absmax = self.absmax.to(device)
            if self.code is not None:
                self.code = self.code.to(device)
            if self.nested:
                self.offset = self.offset.to(device)
                self.state2.to(device)
            return self

    def quantize_fp4(A: torch.Tensor, absmax=None, out=None, blocksize: int = 64,
                     compress_statistics: bool = False, quant_storage=torch.uint8):
        """bitsandbytes.functional.quantize_fp4 equivalent (GPU only)."""
        if not A.is_cuda:
            raise RuntimeError("quantize_fp4 needs a CUDA tensor (there is no CPU quantiser)")
        src = A.contiguous()
        if src.dtype not in (torch.float16, torch.bfloat16, torch.float32):
            src = src.float()
        packed, am = ext.quantize_fp4(src, blocksize)
        code = FP4_CODE.to(A.device)
        if compress_statistics:
            offset = am.mean()
            code2 = create_dynamic_map().to(A.device)
            q, am2 = quantize_blockwise_8bit(am - offset, code2, 256)
            state2 = QuantState(absmax=am2, code=code2, blocksize=256, dtype=torch.float32)
            st = QuantState(absmax=q, shape=A.shape, code=code, blocksize=blocksize,
                            quant_type="fp4", dtype=A.dtype, offset=offset, state2=state2)
        else:
            st = QuantState(absmax=am, shape=A.shape, code=code, blocksize=blocksize,
                            quant_type="fp4", dtype=A.dtype)
        return packed, st

    class Params4bit(nn.Parameter):
        """Minimal bitsandbytes.nn.Params4bit: quantises on the move to a CUDA device."""

        def __new__(cls, data=None, requires_grad=False, quant_state=None, blocksize=64,
                    compress_statistics=False, quant_type="fp4"):
            if data is None:
                data = torch.empty(0)
            self = torch.Tensor._make_subclass(cls, data, requires_grad)
            self.blocksize = blocksize
            self.compress_statistics = compress_statistics
            self.quant_type = quant_type
            self.quant_state = quant_state
            return self

        def _quantize(self, device):
            w = self.data.contiguous().half().cuda(device)  # bitsandbytes < 0.43 quantises from fp16
            packed, st = quantize_fp4(w, blocksize=self.blocksize,
                                      compress_statistics=self.compress_statistics)
            self.data = packed
            self.quant_state = st
            return self

        def cuda(self, device=None):
            if self.quant_state is None and self.data.dtype != torch.uint8:
                return self._quantize(device)
            return self.to(device if device is not None else "cuda")

        def to(self, *args, **kwargs):
            device, dtype, non_blocking, _ = torch._C._nn._parse_to(*args, **kwargs)
            if (device is not None and device.type == "cuda" and self.data.device.type == "cpu"
                    and self.quant_state is None and self.data.dtype != torch.uint8):
                return self._quantize(device)
            if self.quant_state is not None and device is not None:
                self.quant_state.to(device)
            new = Params4bit(super().to(device=device, dtype=None if self.data.dtype == torch.uint8 else dtype,
                                        non_blocking=non_blocking),
                             requires_grad=self.requires_grad, quant_state=self.quant_state,
                             blocksize=self.blocksize, compress_statistics=self.compress_statistics,
                             quant_type=self.quant_type)
            return new

    class Linear4bit(nn.Linear):
        def __init__(self, input_features, output_features, bias=True, compute_dtype=None,
                     compress_statistics=True, quant_type="fp4", device=None):
            super().__init__(input_features, output_features, bias, device)
            self.weight = Params4bit(self.weight.data, requires_grad=False,
                                     compress_statistics=compress_statistics, quant_type=quant_type)
            self.compute_dtype = compute_dtype

        def _apply(self, fn, recurse=True):  # keep Params4bit semantics under .to()/.cuda()
            probe = fn(torch.empty(0))
            if probe.device.type == "cuda" or self.weight.data.dtype == torch.uint8:
                self.weight = self.weight.to(probe.device)
                if self.bias is not None:
                    self.bias = nn.Parameter(fn(self.bias.data), requires_grad=False)
                return self
            return super()._apply(fn, recurse)

        def forward(self, x):
            raise RuntimeError("bnb_compat.Linear4bit is a container; wrap it in TorchFP4Linear")

    class LinearFP4(Linear4bit):
        def __init__(self, input_features, output_features, bias=True, compute_dtype=None,
                     compress_statistics=True, device=None):
            super().__init__(input_features, output_features, bias, compute_dtype,
                             compress_statistics, "fp4", device)

    class _BF:
        QuantState = QuantState
        quantize_fp4 = staticmethod(quantize_fp4)

    BF = _BF()


# ---- on-disk / wire format of a bitsandbytes 4-bit weight ---------------------------
# bitsandbytes (QuantState.as_dict(packed=True) / Linear4bit._save_to_state_dict) stores, next to the packed
# `weight` (uint8): `weight.absmax`, `weight.quant_map` (the 16-entry code), for a double-quantised absmax also
# `weight.nested_absmax` and `weight.nested_quant_map`, and every non-tensor field as JSON in a uint8 tensor
# under `weight.quant_state.bitsandbytes__fp4`.  vLLM's loader reads the same keys
# (vllm/model_executor/model_loader/bitsandbytes_loader.py).  The helpers below write and read that layout for
# any QuantState-like object, so checkpoints saved by bitsandbytes load here and vice versa.
def _pack_dict_to_tensor(d: dict) -> torch.Tensor:
    import json

    return torch.frombuffer(bytearray(json.dumps(d).encode("utf-8")), dtype=torch.uint8).clone()


def _unpack_tensor_to_dict(t: torch.Tensor) -> dict:
    import json

    return json.loads(bytes(t.detach().cpu().to(torch.uint8).numpy().tobytes()).decode("utf-8"))


def quant_state_as_dict(qs, packed: bool = True) -> dict:
    """bitsandbytes QuantState.as_dict: tensors under their own keys, the rest as one packed JSON tensor."""
    d = {"quant_type": qs.quant_type, "absmax": qs.absmax, "blocksize": int(qs.blocksize), "quant_map": qs.code,
         "dtype": str(qs.dtype).replace("torch.", ""), "shape": tuple(int(v) for v in qs.shape)}
    if getattr(qs, "nested", False):
        off = qs.offset
        d.update({"nested_absmax": qs.state2.absmax, "nested_blocksize": int(qs.state2.blocksize),
                  "nested_quant_map": qs.state2.code.clone(),
                  "nested_dtype": str(qs.state2.dtype).replace("torch.", ""),
                  "nested_offset": float(off.item()) if torch.is_tensor(off) else float(off)})
    if not packed:
        return d
    out = {k: v for k, v in d.items() if torch.is_tensor(v)}
    rest = {k: v for k, v in d.items() if not torch.is_tensor(v)}
    out["quant_state.bitsandbytes__" + qs.quant_type] = _pack_dict_to_tensor(rest)
    return out


def quant_state_from_dict(d: dict, device=None):
    """Inverse of quant_state_as_dict (accepts the packed and the unpacked form)."""
    d = dict(d)
    packed_keys = [k for k in d if k.startswith("quant_state.bitsandbytes__")]
    if packed_keys:
        d.update(_unpack_tensor_to_dict(d.pop(packed_keys[0])))
    mv = (lambda t: t.to(device)) if device is not None else (lambda t: t)
    state2, offset = None, None
    if "nested_absmax" in d:
        offset = torch.tensor(float(d["nested_offset"]))
        state2 = QuantState(absmax=mv(d["nested_absmax"]), blocksize=int(d["nested_blocksize"]),
                            code=mv(d["nested_quant_map"]), dtype=getattr(torch, d["nested_dtype"]))
        offset = mv(offset)
    return QuantState(absmax=mv(d["absmax"]), shape=torch.Size(d["shape"]), code=mv(d["quant_map"]),
                      blocksize=int(d["blocksize"]), quant_type=d["quant_type"], dtype=getattr(torch, d["dtype"]),
                      offset=offset, state2=state2)


def make_quantized_linear(weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
                          blocksize: int = 64, compress_statistics: bool = False):
    """Helper for tests/benches: an already-quantised LinearFP4 holding `weight` [out, in] (CUDA)."""
    out_f, in_f = weight.shape
    lin = LinearFP4(in_f, out_f, bias=bias is not None, compress_statistics=compress_statistics)
    packed, st = (BF.quantize_fp4(weight, blocksize=blocksize, compress_statistics=compress_statistics))
    lin.weight = Params4bit(packed, requires_grad=False, quant_state=st, blocksize=blocksize,
                            compress_statistics=compress_statistics, quant_type="fp4")
    if bias is not None:
        lin.bias = nn.Parameter(bias.detach().clone().to(weight.device), requires_grad=False)
    return lin
