import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run with -m gpu on the B200 box")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree shared library (compiled on demand; nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as entry
    entry.build()
    from aru_b200 import engine
    return engine.load_library()
