import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def built():
    """The in-tree libraries (built by __graft_entry__.build(); prebuilt on the GPU box)."""
    import __graft_entry__ as g
    from basic_iterative_solvers_b200 import capi
    if not (os.path.exists(capi.LIB_PATH)):
        g.build()
    return True


@pytest.fixture(scope="session")
def ctx(built):
    """One device context for the whole GPU session.  No fallback: fails without a B200."""
    from basic_iterative_solvers_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()
