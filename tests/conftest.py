import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ops():
    """The product path: reference-named wrappers over torch.ops._C.* (CUDA extension required)."""
    import aphrodite_engine_b200._custom_ops as o
    return o


@pytest.fixture(scope="session")
def cabi():
    from aphrodite_engine_b200 import _native
    return _native.load_c_abi()
