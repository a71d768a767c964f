import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_libs():
    """CPU oracle shared objects (oracle/lib), compiled on demand with make."""
    from oracle import oracle
    oracle.build_libs()
    return oracle


@pytest.fixture(scope="session")
def solver_libs():
    """The sm_100a solver libraries (tpl_b200/lib); nvcc cross-compiles without a GPU."""
    from tpl_b200 import build
    return build.build_zoo()
