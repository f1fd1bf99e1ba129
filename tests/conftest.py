import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_b200():
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a B200 skips the gpu-marked tests instead of erroring in the ctx fixture
    (there is no CPU fallback to run them on)."""
    if _have_b200():
        return
    skip = pytest.mark.skip(reason="needs a B200 (sm_100): the product has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def native_lib():
    """The product library; built in-tree if a toolchain is present and sources are newer."""
    from amatsukaze_b200 import _build
    if _build.needs_build():
        _build.build()
    import amatsukaze_b200 as ab
    return ab.lib()


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle as po
    po.oracle_lib()
    return po


@pytest.fixture(scope="session")
def ctx(native_lib):
    import torch
    import amatsukaze_b200 as ab
    assert torch.cuda.is_available(), "gpu tests need a GPU"
    torch.cuda.set_device(0)
    c = ab.Context(0, torch.cuda.current_stream().cuda_stream)
    yield c
    c.close()
