"""CPU: the C-ABI library loads without a driver and exports every symbol include/transflow_b200.h
declares; the ctypes table covers exactly the header."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "transflow_b200.h")).read()
    return sorted(set(re.findall(r"^TF_API [^;(]*?\b(tf_[a-z0-9_]+)\(", text, re.M)))


def test_library_exports_every_declared_symbol():
    from transflow_b200 import _lib
    names = header_functions()
    assert len(names) >= 40
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    from transflow_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_functions()
    lib = _lib.load()
    assert lib.tf_version() >= 100
    assert ctypes.sizeof(_lib.LayerConfigStruct) == 14 * 4 + 2 * 4 + 8


def test_no_compute_without_gpu_is_an_error_not_a_fallback():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    from transflow_b200 import ops
    with pytest.raises(RuntimeError):
        ops.gray_from_bgr(torch.zeros((4, 4, 3), dtype=torch.uint8))
    import ctypes as C
    from transflow_b200 import _lib
    h = C.c_void_p()
    rc = _lib.load().tf_hs_create(C.byref(h), 32, 32)
    assert rc != 0 and _lib.last_error()
