import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_library():
    """The in-tree C-ABI library (built on demand; nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g

    g.build()
    from mrvi_b200 import _capi

    return _capi.load_library()
