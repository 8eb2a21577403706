"""The C-ABI library loads on a CPU-only box and exports every symbol include/dgp_b200.h declares.
No compute call is made here."""
import ctypes
import re
from pathlib import Path

import pytest

import dynaml_b200
from dynaml_b200 import _capi

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "dgp_b200.h").read_text()


def declared_functions():
    return sorted(set(re.findall(r"\b(dgp_[a-z_0-9]+)\s*\(", HEADER)))


def test_library_is_built_in_tree():
    assert _capi.LIB_PATH.exists(), "run `python -m dynaml_b200.build` (or __graft_entry__.build())"
    assert ROOT in _capi.LIB_PATH.parents


def test_every_declared_symbol_is_exported():
    lib = _capi.load_library()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in dgp_b200.h but not exported"
    assert sorted(_capi.EXPORTS) == names


def test_version_string():
    assert b"sm_100a" in _capi.load_library().dgp_version()


def test_no_cpu_fallback_without_gpu():
    """Creating a context must fail loudly when no sm_100 device is present."""
    lib = _capi.load_library()
    h = ctypes.c_void_p()
    st = lib.dgp_ctx_create(0, ctypes.byref(h))
    if st == _capi.DGP_OK:  # running on a GPU box
        lib.dgp_ctx_destroy(h)
        pytest.skip("a GPU is present")
    assert st == _capi.DGP_ERR_CUDA
    with pytest.raises(_capi.DgpError):
        dynaml_b200.Context(0)


def test_product_never_imports_the_oracle():
    for p in (ROOT / "dynaml_b200").rglob("*.py"):
        assert "oracle" not in p.read_text().replace("no CPU fallback", ""), p
    for p in (ROOT / "dynaml_b200" / "csrc").glob("*"):
        assert "oracle" not in p.read_text(), p
