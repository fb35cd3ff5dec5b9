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


@pytest.fixture(scope="session")
def checker(built):
    """The CPU checker for fresh (non-golden) inputs: the reference's own compiled CVODE/odecommon stack (oracle/_ref, built
    in the container and shipped to the GPU box) whenever it is present, the plain-C restatement only as a stand-in."""
    import oracle

    return oracle.load("ref" if oracle.available("ref") else "port")
