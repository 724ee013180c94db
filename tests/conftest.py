import os
import sys

import pytest

# the oracle's OpenMP team must not spin while pytest-xdist / gloo workers share the cores
os.environ.setdefault("OMP_WAIT_POLICY", "passive")
# the Fortran twins of the library print and exit(1) on an error (the reference's exitt
# behaviour); inside the test runner they report and return instead, so that a failing drop-in
# test fails on its own comparison and the remaining tests still run (fortran_abi.cu: check)
os.environ.setdefault("NEKCEM_B200_TWIN_NO_EXIT", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    # the oracle is test infrastructure: compile it once per session (gcc, seconds)
    from oracle import oracle
    oracle.build()
