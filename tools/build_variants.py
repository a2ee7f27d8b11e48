"""Development: build differently configured copies of libfp4_b200.so for same-box A/B runs.
    python tools/build_variants.py name=-DFLAG1,-DFLAG2 ...     (only csrc/<file> named by src= is rebuilt per variant)
Each variant is linked to torch_bnb_fp4_b200/variants/libfp4_b200.<name>.so; select it with FP4_B200_LIB=<path>."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_bnb_fp4_b200 import build as B  # noqa: E402


def main():
    B.build()  # the default objects
    srcs = ["gemv_stream.cu"]
    variants = []
    for a in sys.argv[1:]:
        if a.startswith("src="):
            srcs = a[4:].split(",")
            continue
        name, _, flags = a.partition("=")
        variants.append((name, [f for f in flags.split(",") if f]))
    outdir = os.path.join(B.HERE, "variants")
    os.makedirs(outdir, exist_ok=True)
    all_objs = sorted(os.path.join(B.OBJ_DIR, f[:-3] + ".o") for f in os.listdir(B.CSRC) if f.endswith(".cu"))

    def one(v):
        name, flags = v
        objs = list(all_objs)
        for s in srcs:
            obj = os.path.join(B.OBJ_DIR, f"{s[:-3]}.{name}.o")
            subprocess.check_call(["nvcc", *B.NVCC_FLAGS, *flags, "-c", os.path.join(B.CSRC, s), "-o", obj])
            objs[objs.index(os.path.join(B.OBJ_DIR, s[:-3] + ".o"))] = obj
        lib = os.path.join(outdir, f"libfp4_b200.{name}.so")
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib, *objs, "-lcudart"])
        print(lib, flush=True)

    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(one, variants))


if __name__ == "__main__":
    main()
