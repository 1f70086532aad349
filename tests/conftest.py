import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the in-tree libraries exist (no-op when they are fresh)."""
    import __graft_entry__ as g

    g.build_cuda()
    from oracle import oracle

    oracle.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o

    o.load()
    return o
