import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Make sure the oracle (test infrastructure) and the engine library exist before any test."""
    from oracle import oracle as orc
    orc.build()
    from qiskit_gym_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    yield
