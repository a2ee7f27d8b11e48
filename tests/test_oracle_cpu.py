"""CPU tests of the oracle itself: conversions against numpy/torch, internal consistency, the
reference's published accuracy band, and the committed golden vectors produced by the UNMODIFIED
reference CUDA extension on a B200 (tests/golden/make_ref_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch

import oracle
from helpers import synth_bytes, synth_quant

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_f16_conversion_matches_numpy():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.standard_normal(4000).astype(np.float32) * s
                        for s in (1e-8, 1e-6, 1e-4, 1e-2, 1, 1e2, 7e4)])
    x = np.concatenate([x, np.array([0, -0.0, 65504, 65519.99, 65520, 2**-24, 2**-25,
                                     2**-25 * 1.0001, 2**-14, np.inf, -np.inf], np.float32)])
    with np.errstate(over="ignore"):
        ref = x.astype(np.float16).view(np.uint16)
    assert np.array_equal(oracle.f32_to_bits(x, oracle.F16), ref)


def test_bf16_conversion_matches_torch():
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.standard_normal(4000).astype(np.float32) * s
                        for s in (1e-30, 1e-8, 1e-3, 1, 1e5, 1e30)])
    ref = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(oracle.f32_to_bits(x, oracle.BF16), ref)


def test_tree_equals_bnb_codebook_bitwise():
    # the tree literals ARE the bitsandbytes code -> both decoders agree bit for bit
    packed, absmax = synth_bytes(64 * 300 + 5, seed=3)
    n = 64 * 300 + 5
    for dt in (oracle.F32, oracle.F16, oracle.BF16):
        a = oracle.dequant_tree(packed, absmax, n, 64, dt)
        b = oracle.dequant_code(packed, absmax, oracle.bnb_code(), n, 64, dt)
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_ref_code_param_differs_from_bnb_as_surveyed():
    # CODE_PARAM is off by -1 / -12 / +2 ulp at nibbles 1 / 4 / 6
    d = oracle.ref_code_param().view(np.int32) - oracle.bnb_code().view(np.int32)
    assert d[1] == -1 and d[4] == -12 and d[6] == 2
    assert all(d[i] == 0 for i in (0, 2, 3, 5, 7))


def test_negative_zero_preserved():
    packed = np.array([0x80, 0x08], np.uint8)  # nibbles 8,0,0,8
    absmax = np.array([0.5], np.float32)
    out = oracle.dequant_tree(packed, absmax, 4, 64, oracle.F32).view(np.uint32)
    assert list(out) == [0x80000000, 0, 0, 0x80000000]


def test_quantize_roundtrip_and_thresholds():
    packed, absmax, w = synth_quant(64 * 512, seed=4)
    d = oracle.dequant_tree(packed, absmax, w.size, 64, oracle.F32)
    # every dequantised value is the nearest-by-threshold code: error bounded by half the largest gap
    blk = np.repeat(absmax, 64)
    assert np.all(np.abs(d - w) <= blk * (1.0 - 0.8333333) + 1e-7)
    # block maxima quantise to +-1 exactly
    for b in range(0, 512, 37):
        seg = slice(b * 64, (b + 1) * 64)
        i = np.argmax(np.abs(w[seg]))
        assert abs(d[seg][i]) == absmax[b]


def test_all_zero_block_quantises_to_zero_nibbles():
    packed, absmax = oracle.quantize(np.zeros(128, np.float32), 64)
    assert np.all(packed == 0) and np.all(absmax == 0)


def test_denest_two_roundings():
    rng = np.random.default_rng(5)
    q = rng.integers(0, 256, 1000, dtype=np.uint8)
    code2 = np.sort(rng.uniform(-1, 1, 256)).astype(np.float32)
    am2 = rng.random(4).astype(np.float32)
    got = oracle.denest(q, code2, am2, 0.0123, 256)
    ref = (code2[q] * am2[np.arange(1000) // 256]).astype(np.float32) + np.float32(0.0123)
    assert np.array_equal(got.view(np.uint32), ref.astype(np.float32).view(np.uint32))


def test_linear_f64_matches_numpy():
    N, K = 48, 256
    packed, absmax, _ = synth_quant(N * K, seed=6)
    x = np.random.default_rng(7).standard_normal((3, K)).astype(np.float32)
    W = oracle.dequant_tree(packed, absmax, N * K, 64, oracle.F32).reshape(N, K).astype(np.float64)
    ref = x.astype(np.float64) @ W.T
    got = oracle.linear_f64(x, packed, absmax, oracle.bnb_code(), None, N, K, 64)
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-12)
    got32 = oracle.linear_f32(x, packed, absmax, oracle.bnb_code(), N, K, 64)
    assert np.allclose(got32, ref, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("dtype,lo,hi", [(oracle.F32, 0.0, 1e-5), (oracle.F16, 2e-4, 6e-3),
                                         (oracle.BF16, 2e-3, 5e-2)])
def test_ref_gemv_emulation_error_magnitudes(dtype, lo, hi):
    # the reference's T-precision accumulation is ~1.5e-2 (bf16) / ~2e-3 (fp16) off
    N, K = 64, 4096
    packed, absmax, _ = synth_quant(N * K, seed=8)
    tdt = {oracle.F32: torch.float32, oracle.F16: torch.float16, oracle.BF16: torch.bfloat16}[dtype]
    x = torch.randn(K, generator=torch.Generator().manual_seed(9)).to(tdt).float().numpy()
    exact = oracle.linear_f64(x[None], packed, absmax, oracle.bnb_code(), None, N, K, 64)[0]
    emu = oracle.gemv_ref_emulate(x, packed, absmax, N, K, 64, dtype)
    err = np.max(np.abs(emu - exact)) / np.max(np.abs(exact))
    assert lo <= err <= hi, err


@pytest.mark.parametrize("tdt", [torch.float32, torch.float16, torch.bfloat16])
def test_readme_accuracy_band(tdt):
    """reference sanity_check.py:130-171 / README.md:90-91,113-115: mean |nn.Linear(x) - fp4(x)| for a
    256->256 layer (default init, seed 10) lies in 0.045-0.065.  Restated on CPU with the oracle
    quantiser (bitsandbytes < 0.43 quantises from fp16, blocksize 64, non-nested)."""
    torch.manual_seed(10)
    lin = torch.nn.Linear(256, 256).to(tdt).requires_grad_(False)
    w16 = lin.weight.data.half().float().numpy().ravel()
    packed, absmax = oracle.quantize(w16, 64)
    odt = {torch.float32: oracle.F32, torch.float16: oracle.F16, torch.bfloat16: oracle.BF16}[tdt]
    for shape in [(1, 256), (2048, 256)]:
        x = torch.randn(*shape).to(tdt)
        ref = torch.nn.functional.linear(x.float(), lin.weight.float(), lin.bias.float()).numpy()
        got = oracle.linear_f64(x.float().numpy(), packed, absmax, oracle.bnb_code(),
                                lin.bias.float().numpy(), 256, 256, 64, odt)
        diff = np.abs(got - ref).mean()
        assert 0.040 <= diff <= 0.070, diff  # single-row cases scatter a little around the band


def _golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "ref_*.npz")))


@pytest.mark.parametrize("path", _golden_files() or [None])
def test_oracle_matches_reference_golden(path):
    """Pins the oracle against outputs of the unmodified reference CUDA extension on B200."""
    if path is None:
        pytest.skip("golden vectors not generated yet (tests/golden/make_ref_golden.py)")
    g = np.load(path)
    n, bs = int(g["n"]), int(g["blocksize"])
    packed, absmax = g["packed"], g["absmax"]
    for name, odt in (("f16", oracle.F16), ("bf16", oracle.BF16), ("f32", oracle.F32)):
        tree = oracle.dequant_tree(packed, absmax, n, bs, odt)
        assert np.array_equal(tree.view(np.uint8), g[f"ref_tree_{name}"].ravel().view(np.uint8)), name
        # the reference's codebook op uses its own CODE_PARAM (it ignores the tensor it is given)
        cb = oracle.dequant_code(packed, absmax, oracle.ref_code_param(), n, bs, odt)
        # (the reference builds with --use_fast_math => FMUL.FTZ; the fixtures avoid fp32 denormals)
        assert np.array_equal(cb.view(np.uint8), g[f"ref_codebook_{name}"].ravel().view(np.uint8)), name
    # GEMV: tolerance (the reference's arithmetic is not IEEE-reproducible)
    N, K = int(g["N"]), int(g["K"])
    for name, odt in (("f16", oracle.F16), ("bf16", oracle.BF16), ("f32", oracle.F32)):
        x = oracle.bits_to_f32(g[f"x_{name}"], odt)
        ref_y = oracle.bits_to_f32(g[f"ref_gemv_{name}"], odt)
        emu = oracle.gemv_ref_emulate(x, packed[: N * K // 2], absmax, N, K, bs, odt)
        exact = oracle.linear_f64(x[None], packed[: N * K // 2], absmax, oracle.ref_code_param(),
                                  None, N, K, bs)[0]
        scale = np.max(np.abs(exact))
        tol_emu = {"f16": 2e-3, "bf16": 1.6e-2, "f32": 1e-5}[name]   # emulation vs real kernel
        tol_exact = {"f16": 1e-2, "bf16": 5e-2, "f32": 1e-5}[name]   # real kernel vs fp64 truth
        assert np.max(np.abs(emu - ref_y)) / scale <= tol_emu, name
        assert np.max(np.abs(ref_y - exact)) / scale <= tol_exact, name
        if name == "bf16":
            # measured on B200: the N4 emulation reproduces the reference's bf16 GEMV bit for bit
            assert np.array_equal(emu, ref_y)
