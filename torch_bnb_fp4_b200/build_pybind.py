"""In-tree build of the compiled `torch_bnb_fp4_ext` pybind module (csrc_torch/torch_fp4.cpp) over libfp4_b200.so.
The module lands in torch_bnb_fp4_b200/pybind/torch_bnb_fp4_ext*.so; putting that directory on sys.path makes
`import torch_bnb_fp4_ext` resolve to it, which is what the reference's own torch_bnb_fp4/__init__.py:11-18 imports."""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc_torch", "torch_fp4.cpp")
OUT_DIR = os.path.join(HERE, "pybind")
NAME = "torch_bnb_fp4_ext"


def module_path() -> str:
    return os.path.join(OUT_DIR, NAME + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force: bool = False, verbose: bool = False) -> str:
    import importlib.util  # (loaded by path: importing the package needs the library this builds)
    spec = importlib.util.spec_from_file_location("fp4_b200_build", os.path.join(HERE, "build.py"))
    core = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(core)
    core.build()
    out = module_path()
    deps = [SRC, os.path.join(HERE, "..", "include", "fp4_b200.h")]
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(d) for d in deps):
        return out
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT_DIR, exist_ok=True)
    inc = [f"-I{p}" for p in ce.include_paths(device_type="cuda")] + [f"-I{os.path.join(HERE, '..', 'include')}",
                                                                f"-I{sysconfig.get_paths()['include']}"]
    libdirs = [f"-L{p}" for p in ce.library_paths(device_type="cuda")] + [f"-L{HERE}"]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, SRC, *inc, *libdirs,
           f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
           "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart",
           "-l:libfp4_b200.so", "-Wl,-rpath,$ORIGIN/..",
           *[f"-Wl,-rpath,{p}" for p in ce.library_paths(device_type="cuda")]]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
