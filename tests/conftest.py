import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ab():
    """The ctypes binding with the device runtime initialised (GPU tests only)."""
    import amrex_b200
    amrex_b200.init(0)
    yield amrex_b200
