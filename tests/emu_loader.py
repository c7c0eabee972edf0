"""TEST INFRASTRUCTURE: binds tests/emu/libblobs_b200_emu.so - the CUDA sources of blobs_b200/csrc compiled by g++ against
a fiber-based stand-in for the CUDA runtime (tests/emu/include/cuda_runtime.h) - so that the parity tests can exercise
the kernels' logic on a machine without a GPU. The blobs_b200 package itself never loads this library: the swap happens
here, inside a context manager, by replacing the handle that blobs_b200._lib.load() caches."""
import contextlib
import ctypes as C
import os
import subprocess

from blobs_b200 import _abi as A
from blobs_b200 import _lib as L

EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
EMU_PATH = os.path.join(EMU_DIR, "libblobs_b200_emu.so")
_emu = None


def build():
    subprocess.run(["make", "-s", "-C", EMU_DIR], check=True)


def load_emu():
    global _emu
    if _emu is None:
        build()
        # BLOBS_TEST_EMU_LIB: an alternative build of the same thing, e.g. one compiled with -fsanitize=address
        lib = C.CDLL(os.environ.get("BLOBS_TEST_EMU_LIB", EMU_PATH))
        for name, (res, args) in L.SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        assert lib.blobs_abi_version() == A.ABI_VERSION
        _emu = lib
    return _emu


@contextlib.contextmanager
def emulated():
    """Inside this block blobs_b200.World(...) is backed by the host-compiled build of the kernels."""
    prev = L._lib
    L._lib = load_emu()
    try:
        yield
    finally:
        L._lib = prev
