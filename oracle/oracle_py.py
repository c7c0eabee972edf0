"""TEST INFRASTRUCTURE — ctypes wrapper over oracle/liboracle.so (the CPU restatement of the
reference's Physics::step). Mirrors the method names of blobs_b200.world.World so parity tests can
drive both with the same code. Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs may import this module; the product package never does."""
import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "liboracle.so")

# load the struct layouts without importing the product package (which would dlopen the CUDA library)
_spec = importlib.util.spec_from_file_location("_blobs_abi_for_oracle", os.path.join(_REPO, "blobs_b200", "_abi.py"))
A = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(A)

_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    vp, u64, f32, i32, dbl, sz = C.c_void_p, C.c_uint64, C.c_float, C.c_int, C.c_double, C.c_size_t
    sig = {
        "orc_new": (vp, [f32, f32, i32]), "orc_free": (None, [vp]), "orc_last_error": (C.c_char_p, [vp]),
        "orc_set_param": (i32, [vp, i32, dbl]), "orc_get_param": (dbl, [vp, i32]),
        "orc_body_insert_many": (i32, [vp, sz, vp, vp]), "orc_collider_insert_many": (i32, [vp, sz, vp, vp, vp]),
        "orc_body_remove": (i32, [vp, u64]), "orc_collider_remove": (i32, [vp, u64]), "orc_reset": (i32, [vp]),
        "orc_spring_insert": (i32, [vp, u64, u64, f32, f32, f32, C.POINTER(u64)]), "orc_spring_remove": (i32, [vp, u64]),
        "orc_joint_insert": (i32, [vp, u64, u64, A.Vec2, A.Vec2, f32, C.POINTER(u64)]), "orc_joint_remove": (i32, [vp, u64]),
        "orc_constraint_push": (i32, [vp, A.Vec2, f32]), "orc_constraint_clear": (i32, [vp]),
        "orc_step": (i32, [vp, dbl]), "orc_fixed_step": (i32, [vp, dbl, C.POINTER(i32)]),
        "orc_step_n_timed": (i32, [vp, dbl, C.c_uint32, C.POINTER(dbl)]),
        "orc_body_slots": (u64, [vp]), "orc_collider_slots": (u64, [vp]), "orc_body_count": (u64, [vp]), "orc_collider_count": (u64, [vp]),
        "orc_download_bodies": (i32, [vp, vp, vp, sz]), "orc_download_colliders": (i32, [vp, vp, vp, sz]),
        "orc_body_get": (i32, [vp, u64, vp]), "orc_body_set": (i32, [vp, u64, vp, C.c_uint32]),
        "orc_apply_forces": (i32, [vp, vp, sz]), "orc_body_translate": (i32, [vp, u64, A.Vec2]),
        "orc_body_colliders": (sz, [vp, u64, vp, sz]), "orc_download_cell_coords": (i32, [vp, vp, vp, sz]),
        "orc_pairs_count": (u64, [vp]), "orc_substeps_recorded": (u64, [vp]), "orc_pairs_drain": (i32, [vp, vp, vp, vp]),
        "orc_events_count": (u64, [vp]), "orc_events_drain": (i32, [vp, vp]),
        "orc_collisions_total": (u64, [vp]), "orc_coincident_total": (u64, [vp]),
        "orc_sh_new": (vp, [f32]), "orc_sh_free": (None, [vp]), "orc_sh_insert": (u64, [vp, f32, f32, f32]),
        "orc_sh_insert_with_id": (None, [vp, u64, f32, f32, f32]), "orc_sh_remove": (i32, [vp, u64]),
        "orc_sh_move_point": (i32, [vp, u64, f32, f32]), "orc_sh_next_id": (u64, [vp]),
        "orc_sh_cell_coords": (None, [vp, f32, f32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
        "orc_sh_cell_population": (sz, [vp, C.c_int32, C.c_int32]), "orc_sh_point": (i32, [vp, u64, vp]),
        "orc_sh_query": (sz, [vp, f32, f32, f32, vp, vp, sz]),
        "orc_body_transform": (None, [f32, f32, f32, C.POINTER(A.Affine2)]),
        "orc_affine_mul": (None, [C.POINTER(A.Affine2), C.POINTER(A.Affine2), C.POINTER(A.Affine2)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class OraclePanic(RuntimeError):
    pass


def _split_pairs(a, b, sub_end):
    out = []
    lo = 0
    for e in sub_end:
        e = int(e)
        seg = np.stack([a[lo:e], b[lo:e]], axis=1).astype(np.uint32)
        if len(seg):
            seg = seg[np.lexsort((seg[:, 1], seg[:, 0]))]
        out.append(seg)
        lo = e
    return out


class OracleWorld:
    """CPU oracle world with the same Python surface as blobs_b200.World."""

    def __init__(self, gravity=(0.0, 0.0), use_spatial_hash=False, grid_pairs=False, maintain_spatial_hash=True, record_events=True):
        self._lib = load()
        self._h = C.c_void_p(self._lib.orc_new(gravity[0], gravity[1], int(use_spatial_hash)))
        self._lib.orc_set_param(self._h, 100, float(maintain_spatial_hash))
        self._lib.orc_set_param(self._h, 101, float(record_events))
        self._lib.orc_set_param(self._h, 102, float(grid_pairs))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.orc_free(self._h)
            self._h = None

    __del__ = close

    def _ck(self, rc):
        if rc:
            raise OraclePanic((self._lib.orc_last_error(self._h) or b"").decode())

    def set_param(self, pid, v):
        self._ck(self._lib.orc_set_param(self._h, pid, float(v)))

    def get_param(self, pid):
        return self._lib.orc_get_param(self._h, pid)

    def reset(self):
        self._ck(self._lib.orc_reset(self._h))

    def insert_bodies(self, descs):
        descs = np.ascontiguousarray(descs, dtype=A.BODY_DESC)
        out = np.zeros(len(descs), dtype=np.uint64)
        self._ck(self._lib.orc_body_insert_many(self._h, len(descs), A.ptr(descs), A.ptr(out)))
        return out

    def insert_colliders(self, descs, parents):
        descs = np.ascontiguousarray(descs, dtype=A.COLLIDER_DESC)
        parents = np.ascontiguousarray(parents, dtype=np.uint64)
        out = np.zeros(len(descs), dtype=np.uint64)
        self._ck(self._lib.orc_collider_insert_many(self._h, len(descs), A.ptr(descs), A.ptr(parents), A.ptr(out)))
        return out

    def remove_body(self, h):
        self._ck(self._lib.orc_body_remove(self._h, int(h)))

    def remove_collider(self, h):
        self._ck(self._lib.orc_collider_remove(self._h, int(h)))

    def body_get(self, h):
        st = np.zeros(1, dtype=A.BODY_STATE)
        if self._lib.orc_body_get(self._h, int(h), A.ptr(st)):
            raise KeyError(h)
        return st[0]

    def body_set(self, h, state, mask):
        st = np.ascontiguousarray(np.asarray(state, dtype=A.BODY_STATE).reshape(1))
        if self._lib.orc_body_set(self._h, int(h), A.ptr(st), mask):
            raise KeyError(h)

    def body_translate(self, h, off):
        self._lib.orc_body_translate(self._h, int(h), A.Vec2(*off))

    def body_apply_force(self, h, f):
        # RigidBody::apply_force (rigid_body.rs:155-160)
        st = self.body_get(h).copy()
        if st["body_type"] != A.BODY_STATIC:
            m = np.float32(st["calculated_mass"])
            st["acceleration"]["x"] = np.float32(st["acceleration"]["x"]) + np.float32(f[0]) / m
            st["acceleration"]["y"] = np.float32(st["acceleration"]["y"]) + np.float32(f[1]) / m
            self.body_set(h, st, A.BODY_ACCELERATION)

    def body_colliders(self, h):
        out = np.zeros(64, dtype=np.uint64)
        n = self._lib.orc_body_colliders(self._h, int(h), A.ptr(out), len(out))
        return out[:n]

    def spring_insert(self, a, b, rest, k, c):
        out = C.c_uint64()
        self._ck(self._lib.orc_spring_insert(self._h, int(a), int(b), rest, k, c, C.byref(out)))
        return out.value

    def spring_remove(self, h):
        self._lib.orc_spring_remove(self._h, int(h))

    def joint_insert(self, a, b, anchor_a=(0.0, 0.0), anchor_b=(0.0, 0.0), distance=float("nan")):
        out = C.c_uint64()
        self._ck(self._lib.orc_joint_insert(self._h, int(a), int(b), A.Vec2(*anchor_a), A.Vec2(*anchor_b), distance, C.byref(out)))
        return out.value

    def joint_remove(self, h):
        self._lib.orc_joint_remove(self._h, int(h))

    def constraint_push(self, pos, radius):
        self._lib.orc_constraint_push(self._h, A.Vec2(*pos), radius)

    def constraint_clear(self):
        self._lib.orc_constraint_clear(self._h)

    def step(self, delta=1.0 / 60.0, n=1):
        for _ in range(n):
            self._ck(self._lib.orc_step(self._h, delta))
        return {"collisions": self._lib.orc_collisions_total(self._h), "coincident_pairs": self._lib.orc_coincident_total(self._h)}

    def step_n_timed(self, delta, n):
        secs = C.c_double()
        self._ck(self._lib.orc_step_n_timed(self._h, delta, n, C.byref(secs)))
        return secs.value

    def fixed_step(self, frame_time):
        n = C.c_int()
        self._ck(self._lib.orc_fixed_step(self._h, frame_time, C.byref(n)))
        return {"steps_run": n.value}

    def body_slots(self):
        return self._lib.orc_body_slots(self._h)

    def collider_slots(self):
        return self._lib.orc_collider_slots(self._h)

    def body_count(self):
        return self._lib.orc_body_count(self._h)

    def collider_count(self):
        return self._lib.orc_collider_count(self._h)

    def download_bodies(self):
        n = self.body_slots()
        st = np.zeros(n, dtype=A.BODY_STATE)
        hd = np.zeros(n, dtype=np.uint64)
        self._lib.orc_download_bodies(self._h, A.ptr(st), A.ptr(hd), n)
        return st, hd

    def download_colliders(self):
        n = self.collider_slots()
        st = np.zeros(n, dtype=A.COLLIDER_STATE)
        hd = np.zeros(n, dtype=np.uint64)
        self._lib.orc_download_colliders(self._h, A.ptr(st), A.ptr(hd), n)
        return st, hd

    def read_positions(self):
        st, _ = self.download_bodies()
        return np.stack([st["position"]["x"], st["position"]["y"]], axis=1)

    def apply_forces(self, f):
        f = np.ascontiguousarray(f, dtype=np.float32)
        self._lib.orc_apply_forces(self._h, A.ptr(f), f.shape[0])

    def cell_coords(self):
        n = self.collider_slots()
        cx = np.zeros(n, dtype=np.int32)
        cy = np.zeros(n, dtype=np.int32)
        self._lib.orc_download_cell_coords(self._h, A.ptr(cx), A.ptr(cy), n)
        return cx, cy

    def record_contacts(self, mode, capacity=0):
        self.pairs_drain()
        self.events_drain()

    def pairs_drain(self):
        n = self._lib.orc_pairs_count(self._h)
        ns = self._lib.orc_substeps_recorded(self._h)
        a = np.zeros(max(n, 1), dtype=np.uint32)
        b = np.zeros(max(n, 1), dtype=np.uint32)
        se = np.zeros(max(ns, 1), dtype=np.uint64)
        self._lib.orc_pairs_drain(self._h, A.ptr(a), A.ptr(b), A.ptr(se))
        return _split_pairs(a[:n], b[:n], se[:ns])

    def events_drain(self):
        n = self._lib.orc_events_count(self._h)
        ev = np.zeros(max(n, 1), dtype=A.COLLISION_EVENT)
        self._lib.orc_events_drain(self._h, A.ptr(ev))
        return ev[:n]

    def coincident_total(self):
        return self._lib.orc_coincident_total(self._h)


class OracleSpatialHash:
    """blobs/src/spatial.rs SpatialHash, CPU restatement."""

    def __init__(self, cell_size):
        self._lib = load()
        self._h = C.c_void_p(self._lib.orc_sh_new(cell_size))

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.orc_sh_free(self._h)
            self._h = None

    def insert(self, pos, radius):
        return self._lib.orc_sh_insert(self._h, pos[0], pos[1], radius)

    def insert_with_id(self, pid, pos, radius):
        self._lib.orc_sh_insert_with_id(self._h, pid, pos[0], pos[1], radius)

    def remove(self, pid):
        return bool(self._lib.orc_sh_remove(self._h, pid))

    def move_point(self, pid, off):
        return bool(self._lib.orc_sh_move_point(self._h, pid, off[0], off[1]))

    @property
    def next_id(self):
        return self._lib.orc_sh_next_id(self._h)

    def get_cell_coords(self, pos):
        cx, cy = C.c_int32(), C.c_int32()
        self._lib.orc_sh_cell_coords(self._h, pos[0], pos[1], C.byref(cx), C.byref(cy))
        return cx.value, cy.value

    def cell_population(self, cell):
        return self._lib.orc_sh_cell_population(self._h, cell[0], cell[1])

    def point(self, pid):
        xyr = np.zeros(3, dtype=np.float32)
        if not self._lib.orc_sh_point(self._h, pid, A.ptr(xyr)):
            return None
        return xyr

    def query(self, pos, radius):
        ids = np.zeros(4096, dtype=np.uint64)
        xyr = np.zeros((4096, 3), dtype=np.float32)
        n = self._lib.orc_sh_query(self._h, pos[0], pos[1], radius, A.ptr(ids), A.ptr(xyr), 4096)
        return [(int(ids[i]), tuple(float(v) for v in xyr[i])) for i in range(n)]


def body_transform(rot, pos):
    out = A.Affine2()
    load().orc_body_transform(rot, pos[0], pos[1], C.byref(out))
    return out


def affine_mul(a, b):
    out = A.Affine2()
    load().orc_affine_mul(C.byref(a), C.byref(b), C.byref(out))
    return out
