"""HF wiring: a transformers Mistral model (tiny, random init - there is no network for
checkpoints) converted with recursively_replace_with_fp4_linear, as examples/speed_test_mistral_7b.py of the
reference does with the real one (lm_head ignored, reference __init__.py:788), then run for prefill and for
token-by-token decode with a KV cache."""
import copy

import pytest
import torch

import torch_bnb_fp4

pytestmark = pytest.mark.gpu


def _tiny_mistral(cuda):
    transformers = pytest.importorskip("transformers")
    cfg = transformers.MistralConfig(hidden_size=512, intermediate_size=1024, num_hidden_layers=2,
                                     num_attention_heads=8, num_key_value_heads=2, vocab_size=384,
                                     max_position_embeddings=128, sliding_window=None)
    torch.manual_seed(0)
    return transformers.MistralForCausalLM(cfg).to(cuda).half().eval()


def test_hf_mistral_prefill_and_decode(cuda):
    dense = _tiny_mistral(cuda)
    fp4 = torch_bnb_fp4.recursively_replace_with_fp4_linear(copy.deepcopy(dense), as_dtype=torch.float16)
    layer = fp4.model.layers[0]
    assert type(layer.self_attn.q_proj).__name__ == "_GroupMember"          # q/k/v share a launch
    assert isinstance(layer.self_attn.o_proj, torch_bnb_fp4.TorchFP4Linear)
    assert type(layer.mlp.gate_proj).__name__ == "_GroupMember"
    assert isinstance(fp4.lm_head, torch.nn.Linear) and not isinstance(fp4.lm_head, torch_bnb_fp4.TorchFP4Linear)
    ids = torch.randint(0, 384, (1, 24), device=cuda, generator=torch.Generator(device=cuda).manual_seed(1))
    with torch.no_grad():
        ref = dense(ids).logits.float()
        got = fp4(ids).logits.float()
        cos = torch.nn.functional.cosine_similarity(ref.flatten(), got.flatten(), dim=0).item()
        assert cos >= 0.93, cos  # FP4 weights (random init, 2 layers): close to, not equal to, the dense model
        # decode: feed the same tokens one by one with a KV cache (GEMV path); must agree with its own prefill
        # (GEMM / dequant path) on every position
        past, outs = None, []
        for t in range(ids.shape[1]):
            o = fp4(ids[:, t:t + 1], past_key_values=past, use_cache=True)
            past = o.past_key_values
            outs.append(o.logits.float())
        dec = torch.cat(outs, dim=1)
        err = ((dec - got).abs().max() / got.abs().max()).item()
        assert err <= 3e-2, err


def test_hf_mistral_fused_mlp(cuda):
    from torch_bnb_fp4_b200._lib import lib
    dense = _tiny_mistral(cuda)
    plain = torch_bnb_fp4.recursively_replace_with_fp4_linear(copy.deepcopy(dense), as_dtype=torch.float16)
    fused = torch_bnb_fp4.recursively_replace_with_fp4_linear(copy.deepcopy(dense), as_dtype=torch.float16)
    assert torch_bnb_fp4.fuse_gated_mlps(fused) == 2  # transformers' SiLUActivation is recognised
    ids = torch.randint(0, 384, (1, 1), device=cuda, generator=torch.Generator(device=cuda).manual_seed(2))
    with torch.no_grad():
        a = plain(ids).logits.float()
        n0 = lib.fp4_b200_launch_count()
        b = fused(ids).logits.float()
        launches = lib.fp4_b200_launch_count() - n0
    assert launches == 2 * (1 + 1 + 2)  # per layer: q/k/v, o, gate/up with the activation, down
    assert ((a - b).abs().max() / a.abs().max()).item() <= 2e-2
