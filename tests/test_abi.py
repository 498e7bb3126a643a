"""CPU: libb200gp.so loads and exports every symbol include/b200gp.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from gpax_b200 import _ffi


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200gp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2gp_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_contract():
    names = declared_symbols()
    for must in ("b2gp_ctx_create", "b2gp_gram", "b2gp_potrf", "b2gp_trsm_lower", "b2gp_gemm_nt",
                 "b2gp_posterior", "b2gp_sparse_posterior", "b2gp_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_ffi.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/b200gp.h but not exported"


def test_binding_covers_the_header():
    assert sorted(_ffi.SIGNATURES) == declared_symbols()
    lib = _ffi.load_library()
    assert lib.b2gp_version() == 100


def test_null_ctx_is_an_argument_error_not_a_crash():
    lib = _ffi.load_library()
    assert lib.b2gp_sync(None) == -1
    assert lib.b2gp_last_error(None) == b"null context"


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="only meaningful on a host without a GPU")
def test_no_silent_cpu_fallback():
    with pytest.raises(_ffi.B200GPError):
        _ffi.Context(0)
