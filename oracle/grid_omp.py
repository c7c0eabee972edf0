"""TEST / BENCH INFRASTRUCTURE: ctypes binding of oracle/libgridomp.so, the all-cores CPU cell-list restatement of
Physics::step for single-collider sphere scenes (see the header of grid_omp.cpp: NOT the reference algorithm, same arithmetic
and summation order, pinned bit-for-bit to the sequential oracle by tests/test_grid_omp.py). Only tests/ and bench.py's
cpu_baseline leg may import this."""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_lib = None


def load():
    global _lib
    if _lib is None:
        path = os.path.join(_DIR, "libgridomp.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", _DIR], check=True)
        lib = C.CDLL(path)
        fp, u8p = C.POINTER(C.c_float), C.POINTER(C.c_uint8)
        lib.blobs_grid_omp_step.restype = C.c_int
        lib.blobs_grid_omp_step.argtypes = [C.c_uint64, fp, fp, fp, fp, fp, fp, fp, fp, fp, u8p, C.c_float, C.c_float, fp, C.c_int, fp, C.c_double,
                                            C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        lib.blobs_grid_omp_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class GridOmpWorld:
    """Holds the state of a blobs_b200.scenes.Scene made of dynamic single-collider default spheres (collider i on body i)."""

    def __init__(self, scene, threads=0):
        b, c = scene.bodies, scene.colliders
        n = len(b)
        assert len(c) == n and np.array_equal(scene.col_parent, np.arange(n)), "one collider per body, in lock step"
        assert len(scene.springs) == 0 and len(scene.joints) == 0
        assert (c["is_sensor"] == 0).all() and (c["memberships"] == 0xFFFFFFFF).all() and (c["filter"] == 0xFFFFFFFF).all()
        assert (c["offset"]["translation"]["x"] == 0).all() and (c["offset"]["translation"]["y"] == 0).all()
        assert (b["body_type"] == 0).all(), "dynamic bodies only"
        xy = lambda v: np.ascontiguousarray(np.stack([v["x"], v["y"]], axis=1), dtype=np.float32)
        self.n = n
        self.pos, self.pos_old, self.acc = xy(b["position"]), xy(b["position_old"]), xy(b["acceleration"])
        self.vel, self.vreq = xy(b["calculated_velocity"]), xy(b["velocity_request"])
        self.has_vreq = np.ascontiguousarray(b["has_velocity_request"], dtype=np.uint8)
        self.snap = xy(c["absolute_transform"]["translation"])
        self.radius = np.ascontiguousarray(c["radius"], dtype=np.float32)
        # calculated_mass = 2 * (2 r): the collider handle is pushed twice, unless a mass override is set
        assert (c["has_mass_override"] == 0).all() if "has_mass_override" in c.dtype.names else True
        self.mass = (np.float32(2.0) * (np.float32(2.0) * self.radius)).astype(np.float32)
        self.gravity_mod = np.ascontiguousarray(b["gravity_mod"], dtype=np.float32)
        self.gravity = tuple(scene.gravity)
        self.constraints = np.ascontiguousarray(np.array(scene.constraints, dtype=np.float32).reshape(-1, 3))
        self.old_dt = np.array([1.0], dtype=np.float32)   # physics.rs:67
        self.substeps = 8
        self.threads = threads
        self.collisions = 0
        self.coincident = 0
        self.seconds = 0.0

    def step(self, delta, n=1):
        lib = load()
        col, coin, secs = C.c_uint64(0), C.c_uint64(0), C.c_double(0.0)
        rc = lib.blobs_grid_omp_step(self.n, _f(self.pos), _f(self.pos_old), _f(self.acc), _f(self.vel), _f(self.snap), _f(self.radius), _f(self.mass),
                                     _f(self.gravity_mod), _f(self.vreq), self.has_vreq.ctypes.data_as(C.POINTER(C.c_uint8)), self.gravity[0], self.gravity[1],
                                     _f(self.constraints), len(self.constraints), _f(self.old_dt), float(delta), self.substeps, int(n), int(self.threads),
                                     C.byref(col), C.byref(coin), C.byref(secs))
        if rc:
            raise RuntimeError(f"blobs_grid_omp_step failed: {rc}")
        self.collisions += col.value
        self.coincident += coin.value
        self.seconds += secs.value
        return {"collisions": col.value, "coincident_pairs": coin.value, "seconds": secs.value}


def max_threads():
    return load().blobs_grid_omp_max_threads()
