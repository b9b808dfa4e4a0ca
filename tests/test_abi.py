"""CPU checks of the C-ABI library: it loads and exports every symbol include/bp_b200.h
declares; without a GPU, context creation fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from ark_bulletproofs_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "bp_b200.h")).read()
    names = set(re.findall(r"\b(bp_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    from ark_bulletproofs_b200 import _lib
    assert names == set(_lib.EXPORTED_SYMBOLS)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.bp_ctx_create(0, 0, ctypes.byref(h))
    assert rc == -6 and not h.value          # BP_ERR_NOGPU
    from ark_bulletproofs_b200 import BpError, Context
    with pytest.raises(BpError):
        Context("secq256k1", 0)
