"""The C-ABI library builds, loads, and exports every symbol that include/vaesne_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "vaesne_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vaesne_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_header_symbols():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as G
    G.build()
    from VAESNe import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    declared = _declared()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_native.EXPORTS) == declared, "ctypes binding table and header disagree"
    lib.vaesne_abi_version.restype = ctypes.c_int
    assert lib.vaesne_abi_version() == 1
    lib.vaesne_is_emulated.restype = ctypes.c_int
    assert lib.vaesne_is_emulated() == 0


def test_product_refuses_cpu_tensors():
    """No CPU fallback: with the real (nvcc) library selected, a CPU tensor raises."""
    import torch
    from VAESNe import _native, _ops
    prev = (_native._lib, _native._emulated)
    try:
        _native.use_library(_native.LIB_PATH)
        with pytest.raises(RuntimeError, match="CUDA tensors"):
            _ops.lin_fwd(torch.zeros(4, 32), torch.zeros(32, 32), torch.zeros(32))
    finally:
        _native._lib, _native._emulated = prev


def test_missing_library_fails_loudly(monkeypatch):
    from VAESNe import _native
    prev = (_native._lib, _native._emulated)
    try:
        _native._lib = None
        monkeypatch.setattr(_native, "LIB_PATH", "/nonexistent/libvaesne_b200.so")
        with pytest.raises(RuntimeError, match="not built"):
            _native.lib()
    finally:
        _native._lib, _native._emulated = prev
