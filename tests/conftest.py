import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "emu: runs the host-compiled build of the CUDA sources (tests/emu) - kernel LOGIC on the CPU")


@pytest.fixture(scope="session", autouse=True)
def _emulated_session():
    """BLOBS_TEST_EMU=1 python -m pytest tests -m gpu ...: run the GPU parity tests on a machine WITHOUT a GPU, against
    tests/emu/libblobs_b200_emu.so (same kernel source, one fiber per CUDA thread; see tests/emu_loader.py). A development
    aid for checking kernel logic before spending GPU time - slow, so pick tests with -k."""
    if os.environ.get("BLOBS_TEST_EMU") != "1":
        yield
        return
    from .emu_loader import emulated

    with emulated():
        yield


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
