"""Host-side mirror of the reference interface: names, enum marshalling, error behaviour."""
import pytest
import torch

import torch_bnb_fp4
import torch_bnb_fp4_ext as ext
from torch_bnb_fp4_b200 import bnb_compat


def test_public_names_of_reference_module():
    # reference torch_bnb_fp4/__init__.py:20-922
    for name in ["ScalarType", "dequantize_fp4", "dequantize_fp4_codebook_invoke_qtype",
                 "dequantize_fp4_codebook_invoke", "gemm_4bit_inference", "gemm_4bit_inference_qtype",
                 "dequantize_fp4_qtype", "QuantData", "TorchFP4Linear", "swap_linear_with_bnb_linear",
                 "check_if_name_contained_in_list", "todevice_if_necessary",
                 "recursively_replace_with_fp4_linear", "T_Model"]:
        assert hasattr(torch_bnb_fp4, name), name
    assert hasattr(torch_bnb_fp4.TorchFP4Linear, "from_linear")


def test_extension_names_of_reference_binding():
    # reference csrc/torch_fp4.cpp:125-139 (+ .export_values())
    for name in ["ScalarType", "bfloat16", "float16", "float32", "dequantize_fp4",
                 "dequantize_fp4_codebook", "gemv_fp4", "qlinear", "qlinear_bias",
                 "qlinear_codebook", "qlinear_codebook_bias"]:
        assert hasattr(ext, name), name
    assert ext.bfloat16 is ext.ScalarType.bfloat16


def test_scalar_type_marshalling():
    S = torch_bnb_fp4.ScalarType
    assert S.from_torch_dtype(torch.bfloat16) is S.bfloat16
    assert S.from_torch_dtype(torch.float16).value is ext.ScalarType.float16
    assert S.from_str("float32") is S.float32
    assert S.bfloat16.torch_dtype is torch.bfloat16
    with pytest.raises(ValueError):
        S.from_torch_dtype(torch.float64)
    with pytest.raises(ValueError):
        S.from_str("int8")
    with pytest.raises(TypeError):
        ext.get_scalar_type(17)


def test_cpu_tensors_are_rejected_like_check_cuda():
    A = torch.zeros(32, 1, dtype=torch.uint8)
    am = torch.ones(1)
    code = torch.tensor(ext.BNB_FP4_CODE)
    with pytest.raises(RuntimeError, match="CUDA"):
        ext.dequantize_fp4(A, am, 64, 8, 8, ext.float16)
    with pytest.raises(RuntimeError, match="CUDA"):
        ext.dequantize_fp4_codebook(A, am, code, 8, 8, 64, 64, ext.bfloat16)
    with pytest.raises(RuntimeError, match="CUDA"):
        ext.gemv_fp4(torch.zeros(1, 64), A, am, code, 64, ext.float32, [1, 64])
    with pytest.raises(RuntimeError, match="CUDA"):
        ext.qlinear(torch.zeros(2, 8), A, am, 8, 8, 64)
    with pytest.raises(RuntimeError):
        bnb_compat.BF.quantize_fp4(torch.randn(64, 64))  # no CPU quantiser


def test_name_filter_and_dynamic_map():
    assert torch_bnb_fp4.check_if_name_contained_in_list("model.lm_head", ["lm_head"])
    assert not torch_bnb_fp4.check_if_name_contained_in_list("q_proj", ["lm_head"])
    code = bnb_compat.create_dynamic_map()
    assert code.numel() == 256 and torch.all(code[1:] >= code[:-1])
    assert code.min() >= -1 and code.max() == 1.0
    v = torch.randn(1000) * 0.01
    q, am2 = bnb_compat.quantize_blockwise_8bit(v, code, 256)
    rec = code[q.long()] * am2[torch.arange(1000) // 256]
    assert (rec - v).abs().max() <= 0.06 * v.abs().max()


def test_surgery_requires_cuda_device():
    with pytest.raises(AssertionError):
        torch_bnb_fp4.recursively_replace_with_fp4_linear(torch.nn.Linear(8, 8), device=torch.device("cpu"))


def test_bnb_4bit_serialisation_round_trip_cpu():
    """QuantState <-> the bitsandbytes 4-bit state_dict keys, plain and nested."""
    import torch
    from torch_bnb_fp4_b200 import bnb_compat as bc

    code = bc.FP4_CODE.clone()
    plain = bc.QuantState(absmax=torch.rand(32), shape=(16, 128), code=code, blocksize=64, quant_type="fp4",
                          dtype=torch.bfloat16)
    d = bc.quant_state_as_dict(plain, packed=True)
    assert set(d) == {"absmax", "quant_map", "quant_state.bitsandbytes__fp4"}
    assert d["quant_state.bitsandbytes__fp4"].dtype == torch.uint8
    meta = bc._unpack_tensor_to_dict(d["quant_state.bitsandbytes__fp4"])
    assert meta == {"quant_type": "fp4", "blocksize": 64, "dtype": "bfloat16", "shape": [16, 128]}
    back = bc.quant_state_from_dict(d)
    assert torch.equal(back.absmax, plain.absmax) and torch.equal(back.code, code)
    assert tuple(back.shape) == (16, 128) and back.blocksize == 64 and back.dtype == torch.bfloat16 and not back.nested

    state2 = bc.QuantState(absmax=torch.rand(1), code=bc.create_dynamic_map(), blocksize=256, dtype=torch.float32)
    nested = bc.QuantState(absmax=torch.randint(0, 256, (32,), dtype=torch.uint8), shape=(16, 128), code=code,
                           blocksize=64, quant_type="fp4", dtype=torch.float16, offset=torch.tensor(0.0371),
                           state2=state2)
    d = bc.quant_state_as_dict(nested, packed=True)
    assert {"nested_absmax", "nested_quant_map"} <= set(d)
    meta = bc._unpack_tensor_to_dict(d["quant_state.bitsandbytes__fp4"])
    assert meta["nested_blocksize"] == 256 and abs(meta["nested_offset"] - 0.0371) < 1e-6
    back = bc.quant_state_from_dict(d)
    assert back.nested and torch.equal(back.absmax, nested.absmax) and back.state2.blocksize == 256
    assert torch.equal(back.state2.code, state2.code) and abs(float(back.offset) - 0.0371) < 1e-6


def test_codebook_identity_accepts_either_sign_of_zero():
    """bitsandbytes' QuantState.code has +0.0 at index 8 (the Python literal -0 is an int); the reference table has
    -0.0.  Both must select the fast kernels (a bitwise compare rejected every real bnb layer)."""
    import torch
    from torch_bnb_fp4_b200 import ext
    ref = torch.tensor(ext.BNB_FP4_CODE, dtype=torch.float32)
    assert ext.code_is_bnb_fp4(ref)
    plus = ref.clone()
    plus[8] = 0.0
    assert not torch.signbit(plus[8]) and ext.code_is_bnb_fp4(plus)
    other = ref.clone()
    other[4] = 0.333333  # the reference's CODE_PARAM quirk: a different table, generic kernel
    assert not ext.code_is_bnb_fp4(other)
    assert not ext.code_is_bnb_fp4(ref[:8].clone())


def test_prefill_dispatch_rule_follows_the_measured_sweep():
    """profiles/r02_gemm_sweep_shapes*.log: which of the two bit-identical prefill paths the dispatcher picks."""
    from torch_bnb_fp4_b200 import _fused_gemm_wins as wins
    # (rows, out_features, in_features) -> fused GEMM faster than dequant + cuBLAS on the B200
    assert wins(16, 4096, 4096) and wins(128, 1024, 4096) and wins(64, 8192, 8192) and wins(128, 28672, 8192)
    assert wins(256, 14336, 4096) and wins(512, 8192, 8192) and wins(512, 28672, 8192) and wins(256, 28672, 4096)
    assert not wins(256, 4096, 4096) and not wins(512, 1024, 4096)          # few row tiles: 0.9x
    assert not wins(16, 4096, 14336) and not wins(512, 4096, 14336)          # few row tiles x long K: 0.6x
    assert wins(64, 28672, 14336)                                            # a full wave of row tiles
    assert not wins(513, 28672, 8192) and not wins(4096, 28672, 8192)        # cuBLAS wins beyond 512 rows
