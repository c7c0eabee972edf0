"""Shared helpers for the parity tests: a backend fixture that yields either the CPU oracle (always) or the
CUDA world (only under -m gpu), behind the same Python surface."""
import numpy as np
import pytest

from blobs_b200 import _abi as A

f32 = np.float32


class Backend:
    def __init__(self, name, factory):
        self.name = name
        self.make = factory
        self.A = A


def backend_params():
    # "emu" = the CUDA sources compiled for the host (tests/emu): kernel logic on the CPU, part of the CPU suite
    return ["oracle", pytest.param("cuda", marks=pytest.mark.gpu), pytest.param("emu", marks=pytest.mark.emu)]


def make_backend(name):
    if name == "oracle":
        from oracle import oracle_py

        return Backend("oracle", lambda gravity=(0.0, 0.0), **kw: oracle_py.OracleWorld(gravity=gravity, **kw))
    import blobs_b200

    if name == "emu":
        from .emu_loader import emulated

        def make(gravity=(0.0, 0.0), **kw):
            with emulated():   # a World keeps the library handle it was created with
                return blobs_b200.World(gravity=gravity, **kw)

        return Backend("emu", make)
    return Backend("cuda", lambda gravity=(0.0, 0.0), **kw: blobs_b200.World(gravity=gravity, **kw))


def sphere(w, pos, r=0.5, col=None, **body_kw):
    """insert_rbd + insert_collider_with_parent for one ball, snapshot = from_translation(position)
    (demo/src/simulation.rs:107,122-146). Returns (body handle, collider handle)."""
    b = A.body_descs(1)
    b["position"]["x"], b["position"]["y"] = pos
    b["position_old"] = b["position"]
    for k, v in body_kw.items():
        if k in ("velocity_request", "acceleration", "calculated_velocity", "position_old"):
            b[k]["x"], b[k]["y"] = v
            if k == "velocity_request":
                b["has_velocity_request"] = 1
        else:
            b[k] = v
    bh = w.insert_bodies(b)
    c = A.collider_descs(1)
    c["radius"] = r
    c["shape_radius"] = r
    c["absolute_transform"]["translation"] = b["position"]
    for k, v in (col or {}).items():
        if k == "offset":
            c["offset"]["translation"]["x"], c["offset"]["translation"]["y"] = v
        else:
            c[k] = v
    ch = w.insert_colliders(c, bh)
    return int(bh[0]), int(ch[0])


def bits(a):
    """float32 array -> uint32 bit patterns with -0.0 folded onto +0.0 (the only tolerated difference in 'bit-exact')."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).copy()
    u[u == 0x80000000] = 0
    return u


def assert_bodies_bit_equal(got, want, fields=("position", "position_old", "calculated_velocity", "acceleration")):
    for f in fields:
        for c in ("x", "y"):
            g, w_ = bits(got[f][c]), bits(want[f][c])
            if not np.array_equal(g, w_):
                bad = np.nonzero(g != w_)[0]
                i = bad[0]
                raise AssertionError(f"{f}.{c}: {len(bad)} of {len(g)} slots differ; first slot {i}: got {got[f][c][i]!r} want {want[f][c][i]!r}")
    for f in ("rotation", "angular_velocity", "calculated_mass", "inertia"):
        if f in fields or f in ("calculated_mass", "inertia"):
            assert np.array_equal(bits(got[f]), bits(want[f])), f
