"""Thin Python handle over the C ABI (include/blobs_b200.h). Used by tests, bench.py and the
`blobs_b200.physics` mirror of the reference's Rust API. All compute happens in libblobs_b200.so."""
import ctypes as C

import numpy as np

from . import _abi as A
from ._lib import load


class BlobsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{A.STATUS_NAMES.get(code, code)}: {msg}")
        self.code = code


def split_pairs(a, b, sub_end):
    """(a, b, running ends) -> list of per-substep (n,2) arrays sorted lexicographically."""
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


class World:
    """One GPU-resident physics world (the reference's `Physics`, physics.rs:3-34)."""

    def __init__(self, gravity=(0.0, 0.0), use_spatial_hash=False, device=-1, body_capacity=0, collider_capacity=0):
        self._lib = load()
        p = A.Params(A.Vec2(*gravity), int(use_spatial_hash), device, body_capacity, collider_capacity)
        h = C.c_void_p()
        rc = self._lib.blobs_world_create(C.byref(p), C.byref(h))
        if rc:
            raise BlobsError(rc, (self._lib.blobs_last_error(None) or b"").decode())
        self._h = h
        self._rec_cap = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.blobs_world_destroy(self._h)
            self._h = None

    __del__ = close

    def _ck(self, rc):
        if rc:
            raise BlobsError(rc, (self._lib.blobs_last_error(self._h) or b"").decode())

    # ---- params
    def set_param(self, pid, v):
        self._ck(self._lib.blobs_world_set_param(self._h, pid, float(v)))

    def get_param(self, pid):
        out = C.c_double()
        self._ck(self._lib.blobs_world_get_param(self._h, pid, C.byref(out)))
        return out.value

    def reset(self):
        self._ck(self._lib.blobs_world_reset(self._h))

    # ---- construction
    def insert_bodies(self, descs):
        descs = np.ascontiguousarray(descs, dtype=A.BODY_DESC)
        out = np.zeros(len(descs), dtype=np.uint64)
        self._ck(self._lib.blobs_body_insert_many(self._h, len(descs), A.ptr(descs), A.ptr(out)))
        return out

    def insert_colliders(self, descs, parents):
        descs = np.ascontiguousarray(descs, dtype=A.COLLIDER_DESC)
        parents = np.ascontiguousarray(parents, dtype=np.uint64)
        assert len(parents) == len(descs)
        out = np.zeros(len(descs), dtype=np.uint64)
        self._ck(self._lib.blobs_collider_insert_many(self._h, len(descs), A.ptr(descs), A.ptr(parents), A.ptr(out)))
        return out

    def remove_body(self, h):
        self._ck(self._lib.blobs_body_remove(self._h, int(h)))

    def remove_collider(self, h):
        self._ck(self._lib.blobs_collider_remove(self._h, int(h)))

    def body_get(self, h):
        st = np.zeros(1, dtype=A.BODY_STATE)
        self._ck(self._lib.blobs_body_get(self._h, int(h), A.ptr(st)))
        return st[0]

    def body_set(self, h, state, mask):
        st = np.ascontiguousarray(np.asarray(state, dtype=A.BODY_STATE).reshape(1))
        self._ck(self._lib.blobs_body_set(self._h, int(h), A.ptr(st), mask))

    def body_translate(self, h, off):
        self._ck(self._lib.blobs_body_translate(self._h, int(h), A.Vec2(*off)))

    def body_apply_force(self, h, f):
        self._ck(self._lib.blobs_body_apply_force(self._h, int(h), A.Vec2(*f)))

    def body_colliders(self, h):
        n = C.c_size_t()
        out = np.zeros(64, dtype=np.uint64)
        self._ck(self._lib.blobs_body_colliders(self._h, int(h), A.ptr(out), len(out), C.byref(n)))
        return out[: n.value]

    def collider_get(self, h):
        st = np.zeros(1, dtype=A.COLLIDER_STATE)
        self._ck(self._lib.blobs_collider_get(self._h, int(h), A.ptr(st)))
        return st[0]

    def spring_insert(self, a, b, rest, k, c):
        out = C.c_uint64()
        self._ck(self._lib.blobs_spring_insert(self._h, int(a), int(b), rest, k, c, C.byref(out)))
        return out.value

    def spring_remove(self, h):
        self._ck(self._lib.blobs_spring_remove(self._h, int(h)))

    def joint_insert(self, a, b, anchor_a=(0.0, 0.0), anchor_b=(0.0, 0.0), distance=float("nan")):
        out = C.c_uint64()
        self._ck(self._lib.blobs_joint_insert(self._h, int(a), int(b), A.Vec2(*anchor_a), A.Vec2(*anchor_b), distance, C.byref(out)))
        return out.value

    def insert_springs(self, a, b, params):
        """bulk springs.insert: a, b handle arrays; params (n,3) = rest_length, stiffness, damping"""
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        params = np.ascontiguousarray(params, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(len(a), dtype=np.uint64)
        self._ck(self._lib.blobs_spring_insert_many(self._h, len(a), A.ptr(a), A.ptr(b), A.ptr(params), A.ptr(out)))
        return out

    def insert_joints(self, a, b, anchors=None, distance=None):
        """bulk create_fixed_joint: anchors (n,4) or None; distance (n,) or None (= measured from current positions)"""
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        anc = None if anchors is None else np.ascontiguousarray(anchors, dtype=np.float32).reshape(-1, 4)
        dst = None if distance is None else np.ascontiguousarray(distance, dtype=np.float32)
        out = np.zeros(len(a), dtype=np.uint64)
        self._ck(self._lib.blobs_joint_insert_many(self._h, len(a), A.ptr(a), A.ptr(b), None if anc is None else A.ptr(anc), None if dst is None else A.ptr(dst), A.ptr(out)))
        return out

    def joint_remove(self, h):
        self._ck(self._lib.blobs_joint_remove(self._h, int(h)))

    def constraint_push(self, pos, radius):
        self._ck(self._lib.blobs_constraint_push(self._h, A.Vec2(*pos), radius))

    def constraint_clear(self):
        self._ck(self._lib.blobs_constraint_clear(self._h))

    # ---- stepping
    def _step_rc(self, rc, st):
        """A strip-decomposed world reports an overflowed message / a neighbour that did not answer through the return code AND
        bits 2 / 3 of nan_detected. Ranks must keep stepping together (one that raised here would leave its peers waiting for its
        messages), so those two are returned in the stats (key "strip_error") instead of raised; everything else raises."""
        d = st.as_dict()
        if rc in (A.ERR_CAPACITY, A.ERR_CUDA) and (d["nan_detected"] & 12):
            d["strip_error"] = (self._lib.blobs_last_error(self._h) or b"").decode()
            return d
        self._ck(rc)
        return d

    def step(self, delta=1.0 / 60.0, n=1):
        st = A.StepStats()
        rc = self._lib.blobs_step(self._h, delta, C.byref(st)) if n == 1 else self._lib.blobs_step_n(self._h, delta, n, C.byref(st))
        return self._step_rc(rc, st)

    def fixed_step(self, frame_time):
        st = A.StepStats()
        return self._step_rc(self._lib.blobs_fixed_step(self._h, frame_time, C.byref(st)), st)

    # ---- state
    def _u64(self, fn):
        out = C.c_uint64()
        self._ck(fn(self._h, C.byref(out)))
        return out.value

    def body_slots(self):
        return self._u64(self._lib.blobs_body_slots)

    def collider_slots(self):
        return self._u64(self._lib.blobs_collider_slots)

    def body_count(self):
        return self._u64(self._lib.blobs_body_count)

    def collider_count(self):
        return self._u64(self._lib.blobs_collider_count)

    def download_bodies(self):
        n = self.body_slots()
        st = np.zeros(n, dtype=A.BODY_STATE)
        hd = np.zeros(n, dtype=np.uint64)
        self._ck(self._lib.blobs_download_bodies(self._h, A.ptr(st), A.ptr(hd), n))
        return st, hd

    def download_colliders(self):
        n = self.collider_slots()
        st = np.zeros(n, dtype=A.COLLIDER_STATE)
        hd = np.zeros(n, dtype=np.uint64)
        self._ck(self._lib.blobs_download_colliders(self._h, A.ptr(st), A.ptr(hd), n))
        return st, hd

    def read_positions(self, out=None):
        n = self.body_slots()
        if out is None:
            out = np.zeros((n, 2), dtype=np.float32)
        self._ck(self._lib.blobs_read_body_positions(self._h, C.c_void_p(out.ctypes.data), n))
        return out

    def read_positions_ptr(self, ptr, n):
        self._ck(self._lib.blobs_read_body_positions(self._h, C.c_void_p(ptr), n))

    def read_velocities(self):
        n = self.body_slots()
        out = np.zeros((n, 2), dtype=np.float32)
        self._ck(self._lib.blobs_read_body_velocities(self._h, A.ptr(out), n))
        return out

    def apply_forces(self, f):
        f = np.ascontiguousarray(f, dtype=np.float32)
        self._ck(self._lib.blobs_apply_forces(self._h, A.ptr(f), f.shape[0]))

    def apply_forces_ptr(self, ptr, n):
        self._ck(self._lib.blobs_apply_forces(self._h, C.c_void_p(ptr), n))

    # pipelined host I/O (include/blobs_b200.h): copies on their own streams, overlapping the neighbouring steps
    def forces_upload_async_ptr(self, ptr, n):
        self._ck(self._lib.blobs_forces_upload_async(self._h, C.c_void_p(ptr), n))

    def apply_forces_uploaded(self):
        self._ck(self._lib.blobs_apply_forces_uploaded(self._h))

    def read_positions_async_ptr(self, ptr, n):
        self._ck(self._lib.blobs_read_body_positions_async(self._h, C.c_void_p(ptr), n))

    def io_sync(self):
        self._ck(self._lib.blobs_io_sync(self._h))

    def cell_coords(self):
        n = self.collider_slots()
        cx = np.zeros(n, dtype=np.int32)
        cy = np.zeros(n, dtype=np.int32)
        self._ck(self._lib.blobs_download_cell_coords(self._h, A.ptr(cx), A.ptr(cy), n))
        return cx, cy

    # ---- contact output
    def record_contacts(self, mode, capacity=1 << 20):
        self._ck(self._lib.blobs_record_contacts(self._h, mode, capacity))
        self._rec_cap = capacity

    def pairs_drain(self):
        """-> list (one entry per substep since the last drain) of sorted (n,2) arrays of (slot_a, slot_b)."""
        cap = self._rec_cap
        a = np.zeros(cap, dtype=np.uint32)
        b = np.zeros(cap, dtype=np.uint32)
        se = np.zeros(8192, dtype=np.uint64)
        n = C.c_size_t()
        ns = C.c_size_t()
        self._ck(self._lib.blobs_pairs_drain(self._h, A.ptr(a), A.ptr(b), cap, C.byref(n), A.ptr(se), len(se), C.byref(ns)))
        return split_pairs(a[: n.value], b[: n.value], se[: ns.value])

    def events_drain(self):
        cap = self._rec_cap
        ev = np.zeros(cap, dtype=A.COLLISION_EVENT)
        n = C.c_size_t()
        self._ck(self._lib.blobs_events_drain(self._h, A.ptr(ev), cap, C.byref(n)))
        return ev[: n.value]

    # ---- multi-GPU strips (config #5)
    @staticmethod
    def strip_unique_id():
        lib = load()
        out = np.zeros(128, dtype=np.uint8)
        rc = lib.blobs_strip_unique_id(A.ptr(out))
        if rc:
            raise BlobsError(rc, (lib.blobs_last_error(None) or b"").decode())
        return out

    def strip_configure(self, rank, nranks, x_lo, x_hi, unique_id=None, ghost_capacity=1 << 16, migrate_capacity=1 << 12):
        uid = None if unique_id is None else np.ascontiguousarray(unique_id, dtype=np.uint8)
        self._ck(self._lib.blobs_strip_configure(self._h, rank, nranks, x_lo, x_hi, None if uid is None else A.ptr(uid), ghost_capacity, migrate_capacity))

    def strip_owned(self):
        n = self.body_slots()
        out = np.zeros(n, dtype=np.uint8)
        self._ck(self._lib.blobs_strip_owned(self._h, A.ptr(out), n))
        return out.astype(bool)

    def read_owned_positions_ptr(self, slots_ptr, xy_ptr, cap):
        n = C.c_size_t()
        self._ck(self._lib.blobs_read_owned_positions(self._h, C.c_void_p(slots_ptr), C.c_void_p(xy_ptr), cap, C.byref(n)))
        return n.value

    def apply_forces_indexed_ptr(self, slots_ptr, fxy_ptr, n):
        self._ck(self._lib.blobs_apply_forces_indexed(self._h, C.c_void_p(slots_ptr), C.c_void_p(fxy_ptr), n))

    # pipelined forms (copies on their own streams; io_sync() completes them)
    def forces_indexed_upload_async_ptr(self, slots_ptr, fxy_ptr, n):
        self._ck(self._lib.blobs_forces_indexed_upload_async(self._h, C.c_void_p(slots_ptr), C.c_void_p(fxy_ptr), n))

    def apply_forces_indexed_uploaded(self):
        self._ck(self._lib.blobs_apply_forces_indexed_uploaded(self._h))

    def read_owned_positions_async_ptr(self, slots_ptr, xy_ptr, n_out_ptr, cap):
        self._ck(self._lib.blobs_read_owned_positions_async(self._h, C.c_void_p(slots_ptr), C.c_void_p(xy_ptr), C.c_void_p(n_out_ptr), cap))

    # ---- introspection
    def kernel_info(self):
        k = A.KernelInfo()
        self._ck(self._lib.blobs_kernel_info(self._h, C.byref(k)))
        return k.as_dict()

    def query_circles(self, centres, radii, flags=0, groups=None, exclude_collider=0, exclude_rigid_body=0, batch_world=0):
        """n circle queries against the live collider snapshots (SpatialHash::query semantics, spatial.rs:155-195, plus the
        QueryFilter the reference left as a stub). Returns (offsets[n+1], hits): hits[offsets[q]:offsets[q+1]] are the collider
        handles of query q in ascending slot order. groups = (memberships, filter) or None."""
        centres = np.ascontiguousarray(centres, dtype=np.float32).reshape(-1, 2)
        n = len(centres)
        radii = np.ascontiguousarray(np.broadcast_to(np.asarray(radii, dtype=np.float32), (n,)))
        f = A.QueryFilter(flags=int(flags), has_groups=int(groups is not None), memberships=int(groups[0]) if groups else 0,
                          filter=int(groups[1]) if groups else 0, exclude_collider=int(exclude_collider),
                          exclude_rigid_body=int(exclude_rigid_body), batch_world=int(batch_world))
        offsets = np.zeros(n + 1, dtype=np.uint64)
        nh = C.c_size_t(0)
        cap = max(16 * n, 64)
        while True:
            hits = np.zeros(cap, dtype=np.uint64)
            rc = self._lib.blobs_query_circles(self._h, n, A.ptr(centres), A.ptr(radii), C.byref(f), A.ptr(offsets), A.ptr(hits), cap, C.byref(nh))
            if rc == A.ERR_CAPACITY and nh.value > cap:
                cap = nh.value
                continue
            self._ck(rc)
            return offsets, hits[: nh.value]

    def debug_data(self):
        """Physics::debug_data (physics.rs:479-481, debug.rs:34-91) in one call: dict of float32 arrays in arena order -
        bodies (n, 6) and colliders (n, 6) as glam Affine2 (x_axis, y_axis, translation), collider_radius (n,),
        joints (n, 4) and springs (n, 4) as (body_a.xy, body_b.xy)."""
        cnt = A.DebugCounts()
        self._ck(self._lib.blobs_debug_counts(self._h, C.byref(cnt)))
        out = {"bodies": np.zeros((cnt.bodies, 6), np.float32), "joints": np.zeros((cnt.joints, 4), np.float32),
               "colliders": np.zeros((cnt.colliders, 6), np.float32), "collider_radius": np.zeros(cnt.colliders, np.float32),
               "springs": np.zeros((cnt.springs, 4), np.float32)}
        self._ck(self._lib.blobs_debug_data(self._h, A.ptr(out["bodies"]), A.ptr(out["joints"]), A.ptr(out["colliders"]),
                                            A.ptr(out["collider_radius"]), A.ptr(out["springs"]), C.byref(cnt)))
        return out

    def profile_enable(self, on=True):
        """0 / False off, 1 / True every kernel class, 2 the dominant kernel only"""
        self._ck(self._lib.blobs_profile_enable(self._h, int(on)))

    def profile_read(self):
        names = ["main", "scan", "scatter", "springs", "joints", "integrate", "other", "strip_pack", "strip_ghost", "nccl_exchange", "crowded", "list_build", "list_decide"]
        ms = np.zeros(len(names), dtype=np.float32)
        nl = np.zeros(len(names), dtype=np.uint64)
        self._ck(self._lib.blobs_profile_read(self._h, A.ptr(ms), A.ptr(nl), len(names)))
        return {k: (float(m), int(n)) for k, m, n in zip(names, ms, nl)}
