import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle module (test infrastructure only)."""
    from oracle import oracle_py

    oracle_py.load()
    return oracle_py


@pytest.fixture(scope="session")
def blobs():
    """The product package; importing it dlopens libblobs_b200.so (no compute without a GPU)."""
    import blobs_b200

    return blobs_b200


@pytest.fixture(scope="session")
def scenes():
    from blobs_b200 import scenes as s

    return s
