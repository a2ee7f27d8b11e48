"""Runs in a SUBPROCESS (tests/test_gpu_reference_module.py): the reference's own, unmodified
torch_bnb_fp4/__init__.py (a git-ignored copy under tests/_ref_module/, made by make_ref_module_copy.py where
/root/reference exists) on top of THIS repo's compiled `torch_bnb_fp4_ext` pybind module
(torch_bnb_fp4_b200/pybind/, csrc_torch/torch_fp4.cpp over libfp4_b200.so), with torch_bnb_fp4_b200.bnb_compat
registered as `bitsandbytes` (not installed in this image).  Prints one JSON object.

What it does is reference sanity_check.py:130-171 (`check`): a 256x256 nn.Linear against its FP4 twin for the
GEMV-3dim, GEMV-2dim and GEMM-3dim inputs, in fp32 / fp16 / bf16, plus the 6-layer MLP of `check_speed`
(:65-122) timed eagerly at batch 1 and 2.  The bnb layers are created the way accelerate's
replace_with_bnb_layers does with the reference's config (sanity_check.py:17-24: fp4, no double quantisation)."""
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch import nn  # noqa: E402

from torch_bnb_fp4_b200 import bnb_compat  # noqa: E402  (loads libfp4_b200.so through ctypes for the quantiser only)

# `bitsandbytes` stand-in, exactly the names torch_bnb_fp4/__init__.py:7-9 imports
bnb = types.ModuleType("bitsandbytes")
bnb.functional = types.ModuleType("bitsandbytes.functional")
bnb.functional.QuantState = bnb_compat.QuantState
bnb.functional.quantize_fp4 = bnb_compat.BF.quantize_fp4
bnb.nn = types.ModuleType("bitsandbytes.nn")
bnb.nn.modules = types.ModuleType("bitsandbytes.nn.modules")
for _n in ("Linear4bit", "LinearFP4", "Params4bit"):
    setattr(bnb.nn, _n, getattr(bnb_compat, _n))
    setattr(bnb.nn.modules, _n, getattr(bnb_compat, _n))
for _m in (bnb, bnb.functional, bnb.nn, bnb.nn.modules):
    sys.modules[_m.__name__] = _m

# the compiled extension and the reference module take precedence over this repo's own packages of the same names
for _k in [k for k in sys.modules if k == "torch_bnb_fp4" or k.startswith("torch_bnb_fp4.") or k == "torch_bnb_fp4_ext"]:
    del sys.modules[_k]
sys.path.insert(0, os.path.join(ROOT, "tests", "_ref_module"))
sys.path.insert(0, os.path.join(ROOT, "torch_bnb_fp4_b200", "pybind"))
import torch_bnb_fp4_ext  # noqa: E402
import torch_bnb_fp4 as ref  # noqa: E402

assert torch_bnb_fp4_ext.__file__.endswith(".so"), torch_bnb_fp4_ext.__file__
assert os.path.join("tests", "_ref_module") in ref.__file__, ref.__file__


def to_bnb(model, dtype):
    """accelerate.utils.bnb.replace_with_bnb_layers(load_in_4bit, fp4, no double quant): nn.Linear -> Linear4bit"""
    for name, child in list(model.named_children()):
        if isinstance(child, nn.Linear):
            new = bnb_compat.Linear4bit(child.in_features, child.out_features, bias=child.bias is not None,
                                        compute_dtype=dtype, compress_statistics=False, quant_type="fp4")
            new.weight = bnb_compat.Params4bit(child.weight.data.clone(), requires_grad=False,
                                               compress_statistics=False, quant_type="fp4").cuda(0)  # quantises
            assert new.weight.data.dtype == torch.uint8 and new.weight.quant_state is not None
            if child.bias is not None:
                new.bias = nn.Parameter(child.bias.data.clone(), requires_grad=False)
            model._modules[name] = new
        else:
            to_bnb(child, dtype)
    return model


class TinyModel(nn.Module):  # sanity_check.py:29-35
    def __init__(self, i, o):
        super().__init__()
        self.in_proj = nn.Linear(i, o)

    def forward(self, x):
        return self.in_proj(x)


class TestModel(nn.Module):  # sanity_check.py:38-50
    def __init__(self, in_dim, hidden, num_hidden, out_dim):
        super().__init__()
        self.in_proj = nn.Linear(in_dim, hidden)
        self.blocks = nn.Sequential(*([nn.GELU(), nn.Linear(hidden, hidden)] * num_hidden))
        self.out_proj = nn.Linear(hidden, out_dim)

    def forward(self, x):
        return self.out_proj(self.blocks(self.in_proj(x)))


def main():
    out = {"ext": torch_bnb_fp4_ext.__file__, "ref_module": ref.__file__, "check": {}, "speed_us": {}}
    for dtype in (torch.float32, torch.float16, torch.bfloat16):
        torch.cuda.manual_seed_all(10)
        torch.manual_seed(10)
        gen = torch.Generator("cuda").manual_seed(10)
        model = TinyModel(256, 256).cuda().type(dtype)
        twin = TinyModel(256, 256).cuda().type(dtype)
        twin.in_proj.weight.data = model.in_proj.weight.data.clone()
        twin.in_proj.bias.data = model.in_proj.bias.data.clone()
        hijack = ref.recursively_replace_with_fp4_linear(to_bnb(twin, dtype), as_dtype=dtype).to("cuda", dtype=dtype)
        assert type(hijack.in_proj).__name__ == "TorchFP4Linear" and type(hijack.in_proj).__module__ == ref.__name__
        ins = {"gemv_3dim": torch.randn(1, 1, 256, generator=gen, device="cuda").type(dtype),
               "gemv_2dim": torch.randn(1, 256, generator=gen, device="cuda").type(dtype),
               "gemm_3dim": torch.randn(1, 2048, 256, generator=gen, device="cuda").type(dtype)}
        res = {}
        with torch.inference_mode():
            for k, x in ins.items():
                y, yq = model(x), hijack(x)
                assert y.shape == yq.shape and yq.dtype == dtype
                res[k] = float((y - yq).abs().mean())
        out["check"][str(dtype)] = res
        # check_speed's model, eager, batch 1 (GEMV path) and 2 (dequant + linear path)
        torch.manual_seed(10)
        mlp = ref.recursively_replace_with_fp4_linear(to_bnb(TestModel(768, 2048, 4, 64).cuda().type(dtype), dtype),
                                                      as_dtype=dtype)
        sp = {}
        with torch.inference_mode():
            for b in (1, 2):
                x = torch.randn(b, 768, device="cuda").type(dtype)
                for _ in range(20):
                    mlp(x)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(200):
                    mlp(x)
                torch.cuda.synchronize()
                sp[f"batch{b}"] = (time.perf_counter() - t0) / 200 * 1e6
        out["speed_us"][str(dtype)] = sp
    print("REFMODULE_JSON " + json.dumps(out))


if __name__ == "__main__":
    main()
