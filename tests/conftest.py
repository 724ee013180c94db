import os
import sys

import pytest

# the oracle's OpenMP team must not spin while pytest-xdist / gloo workers share the cores
os.environ.setdefault("OMP_WAIT_POLICY", "passive")

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
