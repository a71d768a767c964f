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


@pytest.fixture(scope="session")
def cpu_solver(oracle_libs):
    """``cpu_solver(model)`` -> class / factory of the CPU solver the CUDA path is checked against:
    the REAL reference build (oracle/_ref — built in the container by oracle/build_ref.py and
    shipped to the GPU box) when it is there, the C restatement otherwise.  The 7x2 model is
    compared with the strict-IEEE build of the reference (SURVEY.md finding 6)."""
    from oracle import ref

    def get(model):
        Ref = ref.load(model, "strict" if model == "trajectory_tracking_mpc" else "fast")
        if Ref is not None:
            return Ref
        return lambda: oracle_libs.OracleOptim(model)

    get.kind = "reference" if ref.available() else "port"
    return get
