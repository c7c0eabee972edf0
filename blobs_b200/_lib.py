"""Loader for libblobs_b200.so (the CUDA library). There is no CPU fallback: if the shared object
is missing the import fails loudly; if no GPU is present `World()` raises."""
import ctypes as C
import os

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BLOBS_B200_LIBRARY") or os.path.join(_HERE, "libblobs_b200.so")   # (override: A/B of differently compiled builds)

_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/blobs_b200.h declares
SIGNATURES = {
    "blobs_abi_version": (C.c_int32, []),
    "blobs_world_create": (C.c_int32, [C.POINTER(A.Params), C.POINTER(_vp)]),
    "blobs_world_destroy": (C.c_int32, [_vp]),
    "blobs_world_reset": (C.c_int32, [_vp]),
    "blobs_last_error": (C.c_char_p, [_vp]),
    "blobs_world_set_param": (C.c_int32, [_vp, C.c_int32, C.c_double]),
    "blobs_world_get_param": (C.c_int32, [_vp, C.c_int32, C.POINTER(C.c_double)]),
    "blobs_body_insert": (C.c_int32, [_vp, _vp, _u64p]),
    "blobs_body_insert_many": (C.c_int32, [_vp, C.c_size_t, _vp, _vp]),
    "blobs_body_remove": (C.c_int32, [_vp, C.c_uint64]),
    "blobs_body_get": (C.c_int32, [_vp, C.c_uint64, _vp]),
    "blobs_body_set": (C.c_int32, [_vp, C.c_uint64, _vp, C.c_uint32]),
    "blobs_body_count": (C.c_int32, [_vp, _u64p]),
    "blobs_body_translate": (C.c_int32, [_vp, C.c_uint64, A.Vec2]),
    "blobs_body_apply_force": (C.c_int32, [_vp, C.c_uint64, A.Vec2]),
    "blobs_body_colliders": (C.c_int32, [_vp, C.c_uint64, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "blobs_collider_insert": (C.c_int32, [_vp, _vp, C.c_uint64, _u64p]),
    "blobs_collider_insert_many": (C.c_int32, [_vp, C.c_size_t, _vp, _vp, _vp]),
    "blobs_collider_remove": (C.c_int32, [_vp, C.c_uint64]),
    "blobs_collider_get": (C.c_int32, [_vp, C.c_uint64, _vp]),
    "blobs_collider_count": (C.c_int32, [_vp, _u64p]),
    "blobs_spring_insert": (C.c_int32, [_vp, C.c_uint64, C.c_uint64, C.c_float, C.c_float, C.c_float, _u64p]),
    "blobs_spring_remove": (C.c_int32, [_vp, C.c_uint64]),
    "blobs_joint_insert": (C.c_int32, [_vp, C.c_uint64, C.c_uint64, A.Vec2, A.Vec2, C.c_float, _u64p]),
    "blobs_joint_remove": (C.c_int32, [_vp, C.c_uint64]),
    "blobs_spring_insert_many": (C.c_int32, [_vp, C.c_size_t, _vp, _vp, _vp, _vp]),
    "blobs_joint_insert_many": (C.c_int32, [_vp, C.c_size_t, _vp, _vp, _vp, _vp, _vp]),
    "blobs_constraint_push": (C.c_int32, [_vp, A.Vec2, C.c_float]),
    "blobs_constraint_clear": (C.c_int32, [_vp]),
    "blobs_step": (C.c_int32, [_vp, C.c_double, C.POINTER(A.StepStats)]),
    "blobs_fixed_step": (C.c_int32, [_vp, C.c_double, C.POINTER(A.StepStats)]),
    "blobs_step_n": (C.c_int32, [_vp, C.c_double, C.c_uint32, C.POINTER(A.StepStats)]),
    "blobs_body_slots": (C.c_int32, [_vp, _u64p]),
    "blobs_collider_slots": (C.c_int32, [_vp, _u64p]),
    "blobs_download_bodies": (C.c_int32, [_vp, _vp, _vp, C.c_size_t]),
    "blobs_download_colliders": (C.c_int32, [_vp, _vp, _vp, C.c_size_t]),
    "blobs_read_body_positions": (C.c_int32, [_vp, _vp, C.c_size_t]),
    "blobs_read_body_velocities": (C.c_int32, [_vp, _vp, C.c_size_t]),
    "blobs_apply_forces": (C.c_int32, [_vp, _vp, C.c_size_t]),
    "blobs_forces_upload_async": (C.c_int32, [_vp, _vp, C.c_size_t]),
    "blobs_apply_forces_uploaded": (C.c_int32, [_vp]),
    "blobs_read_body_positions_async": (C.c_int32, [_vp, _vp, C.c_size_t]),
    "blobs_io_sync": (C.c_int32, [_vp]),
    "blobs_download_cell_coords": (C.c_int32, [_vp, _vp, _vp, C.c_size_t]),
    "blobs_query_circles": (C.c_int32, [_vp, C.c_size_t, _vp, _vp, C.POINTER(A.QueryFilter), _vp, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "blobs_debug_counts": (C.c_int32, [_vp, C.POINTER(A.DebugCounts)]),
    "blobs_debug_data": (C.c_int32, [_vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(A.DebugCounts)]),
    "blobs_record_contacts": (C.c_int32, [_vp, C.c_int32, C.c_size_t]),
    "blobs_events_drain": (C.c_int32, [_vp, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "blobs_pairs_drain": (C.c_int32, [_vp, _vp, _vp, C.c_size_t, C.POINTER(C.c_size_t), _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "blobs_kernel_info": (C.c_int32, [_vp, C.POINTER(A.KernelInfo)]),
    "blobs_profile_enable": (C.c_int32, [_vp, C.c_int32]),
    "blobs_profile_read": (C.c_int32, [_vp, _vp, _vp, C.c_size_t]),
    "blobs_perf_counter": (None, [C.c_char_p, C.c_uint64]),
    "blobs_perf_counter_inc": (None, [C.c_char_p, C.c_uint64]),
    "blobs_perf_counters_new_frame": (None, [C.c_double]),
    "blobs_perf_counters_reset": (None, []),
    "blobs_perf_counter_get": (C.c_int32, [C.c_char_p, _u64p, C.POINTER(C.c_double)]),
    "blobs_perf_counter_count": (C.c_uint64, []),
    "blobs_perf_counter_at": (C.c_int32, [C.c_uint64, C.c_char_p, C.c_size_t, _u64p, C.POINTER(C.c_double)]),
    "blobs_event_history_len": (C.c_uint64, []),
    "blobs_event_history_get": (C.c_int32, [C.c_uint64, C.POINTER(A.PhysicsEvent)]),
    "blobs_event_history_clear": (None, []),
    "blobs_strip_unique_id": (C.c_int32, [_vp]),
    "blobs_strip_configure": (C.c_int32, [_vp, C.c_int32, C.c_int32, C.c_float, C.c_float, _vp, C.c_uint32, C.c_uint32]),
    "blobs_strip_owned": (C.c_int32, [_vp, _vp, C.c_size_t]),
    "blobs_read_owned_positions": (C.c_int32, [_vp, _vp, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "blobs_forces_indexed_upload_async": (C.c_int32, [_vp, _vp, _vp, C.c_size_t]),
    "blobs_apply_forces_indexed_uploaded": (C.c_int32, [_vp]),
    "blobs_read_owned_positions_async": (C.c_int32, [_vp, _vp, _vp, _vp, C.c_size_t]),
    "blobs_apply_forces_indexed": (C.c_int32, [_vp, _vp, _vp, C.c_size_t]),
}

_lib = None


def load():
    """dlopen the CUDA library and bind every entry point. Raises if the .so is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C blobs_b200/csrc` (or __graft_entry__.build()). "
            "blobs_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.blobs_abi_version() != A.ABI_VERSION:
        raise ImportError("libblobs_b200.so ABI version mismatch")
    _lib = lib
    return lib
