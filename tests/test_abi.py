"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/blobs_b200.h
declares, and the struct layouts used by the Python binding match the header (no compute calls: no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "blobs_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(blobs_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(blobs):
    from blobs_b200 import _lib

    names = _declared()
    assert len(names) >= 40
    lib = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.blobs_abi_version() == 1


def test_struct_layouts_match_header(tmp_path, blobs):
    from blobs_b200 import _abi as A

    prog = tmp_path / "sz.c"
    prog.write_text('#include "blobs_b200.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                    "sizeof(BlobsBodyDesc),sizeof(BlobsBodyState),sizeof(BlobsColliderDesc),sizeof(BlobsColliderState),sizeof(BlobsCollisionEvent),"
                    "sizeof(BlobsStepStats),sizeof(BlobsKernelInfo),offsetof(BlobsBodyState,calculated_mass),offsetof(BlobsColliderDesc,radius));}\n")
    exe = tmp_path / "sz"
    subprocess.check_call([os.environ.get("CC", "gcc"), "-I", os.path.join(REPO, "include"), str(prog), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [A.BODY_DESC.itemsize, A.BODY_STATE.itemsize, A.COLLIDER_DESC.itemsize, A.COLLIDER_STATE.itemsize, A.COLLISION_EVENT.itemsize,
            C.sizeof(A.StepStats), C.sizeof(A.KernelInfo), A.BODY_STATE.fields["calculated_mass"][1], A.COLLIDER_DESC.fields["radius"][1]]
    assert got == want


def test_no_cpu_fallback(blobs):
    """Without a CUDA device world creation must fail loudly (BLOBS_ERR_CUDA), never fall back to the CPU."""
    import torch

    if torch.cuda.is_available():
        return
    try:
        blobs.World()
    except blobs.BlobsError as e:
        assert e.code == blobs.abi.ERR_CUDA
    else:
        raise AssertionError("World() succeeded without a GPU")


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under blobs_b200/ may import, link or dlopen it."""
    pat = re.compile(r"(import\s+oracle|from\s+oracle|oracle_py|liboracle|orc_[a-z_]+\s*\(|blobs_oracle\.hpp)")
    for root, _, files in os.walk(os.path.join(REPO, "blobs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert not pat.search(txt), f"{f} references the oracle"


def test_scene_generators_are_deterministic(scenes):
    a, b = scenes.cfg1(3), scenes.cfg1(3)
    assert a.bodies.tobytes() == b.bodies.tobytes() and a.colliders.tobytes() == b.colliders.tobytes()
    assert scenes.cfg1(4).bodies.tobytes() != a.bodies.tobytes()
    u = scenes.uniform(1, 1, 1000)
    assert u.dtype == np.float32 and (u >= 0).all() and (u < 1).all()
    sc = scenes.cfg4(n_blobs=4, k=16)
    assert sc.n_bodies == 64 and len(sc.joints) == 64 and len(sc.springs) == 64 + 32
