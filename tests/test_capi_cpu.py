"""The C-ABI library loads and exports every symbol include/fp4_b200.h declares; argument
validation that happens before any CUDA call is exercised (no compute without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "fp4_b200.h")).read()
    return sorted(set(re.findall(r"\b(fp4_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = ctypes.CDLL(os.path.join(ROOT, "torch_bnb_fp4_b200", "libfp4_b200.so"))
    names = _declared()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), n


def test_binding_covers_header():
    from torch_bnb_fp4_b200 import _lib

    assert sorted(_lib.EXPORTS) == _declared()
    assert _lib.lib.fp4_b200_abi_version() == 1


def test_status_strings_and_early_validation():
    from torch_bnb_fp4_b200._lib import lib

    assert lib.fp4_b200_status_string(0) == b"ok"
    assert b"NULL" in lib.fp4_b200_status_string(-1)
    # rejected before any CUDA call
    assert lib.fp4_b200_dequantize(None, None, None, None, 16, 64, 0, None) == -1
    assert lib.fp4_b200_gemv(None, None, None, None, None, None, None, 1, 8, 64, 64, 0, 0, None, 0, None) == -1
    assert lib.fp4_b200_gemv(1, 1, 1, None, None, None, 1, 9, 8, 64, 64, 0, 0, None, 0, None) == -6   # batch
    assert lib.fp4_b200_gemv(1, 1, 1, None, None, None, 1, 1, 8, 48, 64, 0, 0, None, 0, None) == -7   # K % 32
    assert lib.fp4_b200_gemv(1, 1, 1, None, None, None, 1, 1, 8, 64, 48, 0, 0, None, 0, None) == -4   # blocksize
    assert lib.fp4_b200_gemv(1, 1, 1, None, None, None, 1, 1, 8, 64, 64, 7, 0, None, 0, None) == -2   # dtype
    assert lib.fp4_b200_gemv_workspace_bytes(4096) == 0  # no kernel needs scratch memory (ABI v1 parameter kept)
    assert lib.fp4_b200_quantize(None, 0, 16, 64, None, None, None) == -1
    assert lib.fp4_b200_dequantize(1, 1, None, 1, 16, 48, 0, None) == -4
    assert lib.fp4_b200_dequantize(1, 1, None, 1, 16, 64, 9, None) == -2
    assert lib.fp4_b200_dequantize(1, 1, None, 1, 0, 64, 0, None) == 0                      # empty


def test_tp_struct_layout_matches_header(tmp_path):
    """ctypes mirror of fp4_b200_tp_t / fp4_b200_nested_t has the C layout (compiled with gcc from the header)."""
    import subprocess

    from torch_bnb_fp4_b200 import _lib

    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "fp4_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(fp4_b200_tp_t),'
                   'offsetof(fp4_b200_tp_t,in_base),offsetof(fp4_b200_tp_t,slot_bytes),offsetof(fp4_b200_tp_t,out_world),'
                   'offsetof(fp4_b200_tp_t,out_peer_base),offsetof(fp4_b200_tp_t,epochs),offsetof(fp4_b200_tp_t,err),'
                   'sizeof(fp4_b200_nested_t));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    T = _lib.TpExchange
    want = [ctypes.sizeof(T), T.in_base.offset, T.slot_bytes.offset, T.out_world.offset, T.out_peer_base.offset,
            T.epochs.offset, T.err.offset, ctypes.sizeof(_lib.Nested)]
    assert got == want


def test_prepared_layer_handle_rejects_null_and_destroy_is_null_safe():
    """fp4_b200_layer_create validates on the host (no GPU needed); a NULL handle is an error, not a crash."""
    from torch_bnb_fp4_b200 import _lib
    lib = _lib.lib
    assert not lib.fp4_b200_layer_create(None, None, None, None, 16, 64, 64, 2, 0)
    assert not lib.fp4_b200_layer_create(4096, 8192, None, None, 0, 64, 64, 2, 0)
    h = lib.fp4_b200_layer_create(4096, 8192, None, None, 16, 64, 64, 2, 0)   # pointers are not dereferenced
    assert h
    lib.fp4_b200_layer_destroy(h)
    lib.fp4_b200_layer_destroy(None)
    assert lib.fp4_b200_layer_gemv(None, None, None, 1, None, 0, None) == -1   # FP4_B200_ERR_NULL

