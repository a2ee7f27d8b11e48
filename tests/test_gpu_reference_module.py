"""Drop-in proof: the reference's OWN, unmodified torch_bnb_fp4/__init__.py runs on this repo's compiled
`torch_bnb_fp4_ext` pybind module (csrc_torch/torch_fp4.cpp over the C-ABI) and lands in the band its sanity check
prints (reference sanity_check.py:130-171, 177-179: 0.045-0.065)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_COPY = os.path.join(ROOT, "tests", "_ref_module", "torch_bnb_fp4", "__init__.py")


@pytest.mark.gpu
def test_unmodified_reference_module_over_compiled_ext(cuda):
    from torch_bnb_fp4_b200 import build_pybind
    if not os.path.exists(REF_COPY):
        pytest.skip("tests/_ref_module/ absent (made by __graft_entry__.build() where /root/reference exists)")
    assert os.path.exists(build_pybind.module_path()), "compiled torch_bnb_fp4_ext missing: run __graft_entry__.build()"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_module_runner.py")], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("REFMODULE_JSON ")][-1]
    res = json.loads(line[len("REFMODULE_JSON "):])
    assert res["ext"].endswith(".so") and "pybind" in res["ext"]
    assert "_ref_module" in res["ref_module"]
    for dtype, checks in res["check"].items():
        for name, diff in checks.items():
            assert 0.040 <= diff <= 0.070, (dtype, name, diff)  # the reference's stated band, with a little slack
    print(json.dumps(res["speed_us"]))


def test_compiled_ext_exports_the_reference_surface():
    """reference csrc/torch_fp4.cpp:125-139: module name, enum with exported values, seven functions, positional
    signatures (checked on the CPU box: import only, no compute)."""
    import importlib.util

    from torch_bnb_fp4_b200 import build_pybind
    path = build_pybind.module_path()
    if not os.path.exists(path):
        pytest.skip("compiled torch_bnb_fp4_ext not built yet")
    spec = importlib.util.spec_from_file_location("torch_bnb_fp4_ext", path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    for name in ("dequantize_fp4", "dequantize_fp4_codebook", "gemv_fp4", "qlinear", "qlinear_bias",
                 "qlinear_codebook", "qlinear_codebook_bias", "ScalarType", "bfloat16", "float16", "float32"):
        assert hasattr(m, name), name
    assert [int(m.ScalarType.float16), int(m.ScalarType.float32), int(m.ScalarType.bfloat16)] == [0, 1, 2]
    doc = m.dequantize_fp4_codebook.__doc__
    assert "arg7: torch_bnb_fp4_ext.ScalarType" in doc and "arg2: torch.Tensor" in doc
    import torch
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        m.dequantize_fp4(torch.zeros(8, dtype=torch.uint8), torch.zeros(1), 64, 1, 16, m.float16)
