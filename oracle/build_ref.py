"""Build the UNMODIFIED reference CUDA extension for sm_100a as a checker / timing baseline.

Compiles the three source files where they lie under /root/reference/csrc (never copied into this
repo) with the reference's own nvcc flags (reference setup.py:34-46) into ``oracle/_ref/`` as the
module ``torch_bnb_fp4_ext_ref``.  ``oracle/_ref/`` is git-ignored but travels to the GPU box with
gpurun, where /root/reference does not exist.  Test infrastructure only.

Usage: python oracle/build_ref.py [--force]
"""
from __future__ import annotations

import glob
import os
import sys

REF_CSRC = "/root/reference/csrc"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "torch_bnb_fp4_ext_ref"
SOURCES = ["gemv_fp4_optimized.cu", "dequant_fp4_optimized.cu", "torch_fp4.cpp"]  # setup.py:71-75

# reference setup.py:34-46
NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
    "-U__CUDA_NO_HALF2_OPERATORS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__",
    "--expt-relaxed-constexpr", "--expt-extended-lambda", "--use_fast_math",
    "--ptxas-options=-allow-expensive-optimizations=true",
    "-gencode", "arch=compute_100a,code=sm_100a",
]


def built_path() -> str | None:
    hits = glob.glob(os.path.join(OUT_DIR, NAME + "*.so"))
    return hits[0] if hits else None


def build(force: bool = False) -> str | None:
    """Returns the path of the built module, or None when the reference sources are absent."""
    srcs = [os.path.join(REF_CSRC, s) for s in SOURCES]
    have = built_path()
    if not all(os.path.exists(s) for s in srcs):
        return have  # GPU box: use the prebuilt file if it travelled
    if have and not force and os.path.getmtime(have) >= max(os.path.getmtime(s) for s in srcs):
        return have
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ.setdefault("MAX_JOBS", "4")
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load

    load(name=NAME, sources=srcs, extra_cflags=["-O3", "-std=c++17"], extra_cuda_cflags=NVCC_FLAGS,
         build_directory=OUT_DIR, verbose=False, is_python_module=False)
    return built_path()


def load_module():
    """Import the prebuilt reference extension (needs CUDA at call time, not at import)."""
    path = built_path()
    if path is None:
        return None
    import importlib.util

    import torch  # noqa: F401  (the extension links against libtorch)

    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
