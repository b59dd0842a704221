import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def lib():
    """The in-tree C-ABI library; built on demand (nvcc cross-compiles without a GPU)."""
    import __graft_entry__

    __graft_entry__.build()
    import samurai_b200

    return samurai_b200


@pytest.fixture(scope="session")
def gpu(lib):
    if not lib.initialize(0):
        pytest.fail("no CUDA device: the gpu tests have no CPU fallback")
    return lib
