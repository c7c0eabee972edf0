"""ctypes / numpy mirrors of the C structs in include/blobs_b200.h.

The same layouts are used by the CUDA library (libblobs_b200.so) and, in tests only, by the CPU
checker, so one scene description can be fed to both.
"""
import ctypes as C

import numpy as np

ABI_VERSION = 1

# BlobsStatus
OK, ERR_STALE_HANDLE, ERR_SAME_BODY, ERR_NAN, ERR_CUDA, ERR_INVALID, ERR_SPATIAL_HASH, ERR_MASS, ERR_DANGLING, ERR_CAPACITY = range(10)
STATUS_NAMES = {
    0: "BLOBS_OK", 1: "BLOBS_ERR_STALE_HANDLE", 2: "BLOBS_ERR_SAME_BODY", 3: "BLOBS_ERR_NAN", 4: "BLOBS_ERR_CUDA",
    5: "BLOBS_ERR_INVALID", 6: "BLOBS_ERR_SPATIAL_HASH", 7: "BLOBS_ERR_MASS", 8: "BLOBS_ERR_DANGLING", 9: "BLOBS_ERR_CAPACITY",
}

# RigidBodyType (rigid_body.rs:221-242)
BODY_DYNAMIC, BODY_STATIC, BODY_KINEMATIC_POSITION, BODY_KINEMATIC_VELOCITY = range(4)

# BlobsParamId
(PARAM_GRAVITY_X, PARAM_GRAVITY_Y, PARAM_SUBSTEPS, PARAM_JOINT_ITERATIONS, PARAM_USE_SPATIAL_HASH, PARAM_COLLISIONS_ENABLED,
 PARAM_ACCUMULATOR, PARAM_TIME, PARAM_OLD_DT, PARAM_CELL_SIZE, PARAM_BROADPHASE_CELL, PARAM_CONTACT_MODE, PARAM_FUSED, PARAM_TUNE, PARAM_BATCH_WORLD, PARAM_GRAPH, PARAM_GRAPH_REPLAYS, PARAM_STRIP_MAX_GHOSTS,
 PARAM_STRIP_MAX_MIGRANTS, PARAM_CROWDED, PARAM_POOL, PARAM_POOL_MIN, PARAM_STRIP_P2P, PARAM_LIST, PARAM_SKIN, PARAM_LIST_REBUILDS,
 PARAM_LIST_SUBSTEPS, PARAM_LIST_ACTIVE) = range(28)

# body field mask
BODY_POSITION = 1 << 0
BODY_POSITION_OLD = 1 << 1
BODY_ACCELERATION = 1 << 2
BODY_VELOCITY_REQUEST = 1 << 3
BODY_CALC_VELOCITY = 1 << 4
BODY_ROTATION = 1 << 5
BODY_ANGULAR_VELOCITY = 1 << 6
BODY_TORQUE = 1 << 7
BODY_MASS = 1 << 8
BODY_INERTIA = 1 << 9
BODY_GRAVITY_MOD = 1 << 10
BODY_TYPE = 1 << 11
BODY_USER_DATA = 1 << 12
BODY_SCALE = 1 << 13
BODY_CENTER_OF_MASS = 1 << 14
BODY_ALL = 0x7FFF

RECORD_OFF, RECORD_PAIRS, RECORD_EVENTS = range(3)

_vec2 = np.dtype([("x", "<f4"), ("y", "<f4")])
_affine = np.dtype([("x_axis", _vec2), ("y_axis", _vec2), ("translation", _vec2)])

BODY_DESC = np.dtype([
    ("position", _vec2), ("position_old", _vec2), ("gravity_mod", "<f4"), ("rotation", "<f4"), ("scale", _vec2),
    ("acceleration", _vec2), ("velocity_request", _vec2), ("calculated_velocity", _vec2),
    ("has_velocity_request", "<i4"), ("body_type", "<u4"), ("user_data_lo", "<u8"), ("user_data_hi", "<u8"),
], align=True)

BODY_STATE = np.dtype([
    ("position", _vec2), ("position_old", _vec2), ("center_of_mass", _vec2), ("scale", _vec2), ("acceleration", _vec2),
    ("velocity_request", _vec2), ("calculated_velocity", _vec2),
    ("calculated_mass", "<f4"), ("gravity_mod", "<f4"), ("rotation", "<f4"), ("angular_velocity", "<f4"),
    ("torque", "<f4"), ("inertia", "<f4"), ("has_velocity_request", "<i4"), ("body_type", "<u4"),
    ("user_data_lo", "<u8"), ("user_data_hi", "<u8"),
], align=True)

COLLIDER_DESC = np.dtype([
    ("offset", _affine), ("absolute_transform", _affine), ("radius", "<f4"), ("mass_override", "<f4"),
    ("shape_radius", "<f4"), ("has_mass_override", "<i4"), ("is_sensor", "<i4"), ("memberships", "<u4"), ("filter", "<u4"),
    ("user_data_lo", "<u8"), ("user_data_hi", "<u8"),
], align=True)

COLLIDER_STATE = np.dtype([("desc", COLLIDER_DESC), ("parent", "<u8")], align=True)

COLLISION_EVENT = np.dtype([("col_handle_a", "<u8"), ("col_handle_b", "<u8"), ("impact_vel_a", _vec2), ("impact_vel_b", _vec2)], align=True)


class Vec2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class Affine2(C.Structure):
    _fields_ = [("x_axis", Vec2), ("y_axis", Vec2), ("translation", Vec2)]


class Params(C.Structure):
    _fields_ = [("gravity", Vec2), ("use_spatial_hash", C.c_int32), ("device", C.c_int32),
                ("body_capacity_hint", C.c_uint32), ("collider_capacity_hint", C.c_uint32)]


class StepStats(C.Structure):
    _fields_ = [("collisions", C.c_uint64), ("coincident_pairs", C.c_uint64), ("events_dropped", C.c_uint64),
                ("nan_detected", C.c_uint32), ("steps_run", C.c_uint32), ("substeps_run", C.c_uint32),
                ("list_overflow", C.c_uint32), ("gpu_ms", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class KernelInfo(C.Structure):
    _fields_ = [("launches", C.c_uint64), ("grid_w", C.c_uint32), ("grid_h", C.c_uint32), ("broadphase_cell", C.c_float),
                ("r_max", C.c_float), ("fused_path", C.c_uint32), ("n_simple_bodies", C.c_uint32),
                ("n_multi_bodies", C.c_uint32), ("n_spring_bodies", C.c_uint32), ("n_islands", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class QueryFilter(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("has_groups", C.c_int32), ("memberships", C.c_uint32), ("filter", C.c_uint32),
                ("exclude_collider", C.c_uint64), ("exclude_rigid_body", C.c_uint64), ("batch_world", C.c_uint32), ("reserved", C.c_uint32)]


QUERY_EXCLUDE_FIXED, QUERY_EXCLUDE_KINEMATIC, QUERY_EXCLUDE_DYNAMIC, QUERY_EXCLUDE_SENSORS, QUERY_EXCLUDE_SOLIDS = 2, 4, 8, 16, 32


class DebugCounts(C.Structure):
    _fields_ = [("bodies", C.c_uint64), ("joints", C.c_uint64), ("colliders", C.c_uint64), ("springs", C.c_uint64)]


class PhysicsEvent(C.Structure):  # events.rs:42-50
    _fields_ = [("real_time", C.c_double), ("unpaused_time", C.c_double), ("position", Vec2), ("has_position", C.c_int32),
                ("severity", C.c_int32), ("col_handle", C.c_uint64), ("rbd_handle", C.c_uint64), ("message", C.c_char * 64)]


SEVERITY_TRACE, SEVERITY_DEBUG, SEVERITY_INFO, SEVERITY_WARN, SEVERITY_ERROR, SEVERITY_CRITICAL = range(6)


def body_descs(n):
    """n default RigidBodyBuilder::new() descriptors (rigid_body.rs:303-318)."""
    d = np.zeros(n, dtype=BODY_DESC)
    d["gravity_mod"] = 1.0
    d["scale"]["x"] = 1.0
    d["scale"]["y"] = 1.0
    d["body_type"] = BODY_DYNAMIC
    return d


def collider_descs(n):
    """n default ColliderBuilder::new() descriptors (collider.rs:212-224)."""
    d = np.zeros(n, dtype=COLLIDER_DESC)
    for k in ("offset", "absolute_transform"):
        d[k]["x_axis"]["x"] = 1.0
        d[k]["y_axis"]["y"] = 1.0
    d["radius"] = 0.5
    d["shape_radius"] = 0.5
    d["memberships"] = 0xFFFFFFFF
    d["filter"] = 0xFFFFFFFF
    return d


def ptr(a, ctype=C.c_void_p):
    return a.ctypes.data_as(ctype)
