import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Native pieces built in-tree (the CUDA library cross-compiles without a GPU)."""
    import __graft_entry__ as g

    g.build()
    return True


@pytest.fixture(scope="session")
def port(built):
    import oracle

    return oracle.load("port")


@pytest.fixture(scope="session")
def ref(built):
    import oracle

    if not oracle.available("ref"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return oracle.load("ref")
