"""Copy the reference's Python module, UNMODIFIED, into the git-ignored tests/_ref_module/ so that the drop-in
test (test_gpu_reference_module.py) can import it on the GPU box, where /root/reference does not exist.  The copy is
test input (like oracle/_ref/), never part of the product, never committed."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/torch_bnb_fp4/__init__.py"
DST = os.path.join(HERE, "_ref_module", "torch_bnb_fp4", "__init__.py")


def make() -> bool:
    if not os.path.exists(SRC):
        return os.path.exists(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copyfile(SRC, DST)
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
