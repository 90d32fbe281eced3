import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vaesne-dev_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="module")
def emu():
    """Route the product's native binding to the CPU emulator build of the same kernel sources."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    from VAESNe import _native
    path = build_emu.build()
    prev = (_native._lib, _native._emulated)
    _native.use_library(path)
    yield _native
    _native._lib, _native._emulated = prev
