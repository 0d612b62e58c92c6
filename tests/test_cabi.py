"""CPU-side checks: the C-ABI library builds, loads and exports every symbol include/clair3_b200.h declares; the
product path fails loudly (no CPU fallback) when no B200 is present."""
import ctypes
import os

import pytest

from clair3_b200 import _ffi


def test_library_exports_every_declared_symbol():
    path = _ffi.LIB_PATH
    _ffi.lib()
    assert os.path.exists(path)
    dll = ctypes.CDLL(path)
    assert len(_ffi.DECLARED_FUNCTIONS) >= 13
    for name in _ffi.DECLARED_FUNCTIONS:
        assert hasattr(dll, name), name


def test_header_has_no_torch_types():
    import re
    src = open(_ffi.HEADER).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)          # declarations only, comments stripped
    assert "torch" not in code.lower() and "at::" not in code and "std::" not in code
    assert 'extern "C"' in code


def test_version_and_error_strings():
    L = _ffi.lib()
    assert b"sm_100a" in _ffi.ffi.string(L.c3b_version())
    assert L.c3b_out_dim(_ffi.ffi.NULL) == -1


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from clair3_b200.model import Clair3_P
    m = Clair3_P(add_indel_length=False, predict=True, input_channels=18)
    with pytest.raises(_ffi.C3BError):
        m.to(torch.device("cpu"))
    with pytest.raises(_ffi.C3BError, match="no CUDA device|no CPU"):
        m.load_state_dict({})
    out = _ffi.ffi.new("c3b_model **")
    assert _ffi.lib().c3b_create(out, 0, 18, 0, 0) != 0
    assert b"no CPU fallback" in _ffi.ffi.string(_ffi.lib().c3b_last_error())


def test_product_path_never_imports_oracle():
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "clair3_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
