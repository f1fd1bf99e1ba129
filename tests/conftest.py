import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


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
