//! Raw bindings to include/blobs_b200.h (what `bindgen` would emit, trimmed to the entry points the shim uses).
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_double, c_float, c_int};

#[repr(C)] pub struct BlobsWorld { _private: [u8; 0] }
pub type BlobsHandle = u64;

#[repr(C)] #[derive(Copy, Clone, Default)] pub struct BlobsVec2 { pub x: c_float, pub y: c_float }
#[repr(C)] #[derive(Copy, Clone, Default)] pub struct BlobsAffine2 { pub x_axis: BlobsVec2, pub y_axis: BlobsVec2, pub translation: BlobsVec2 }

#[repr(C)] pub struct BlobsParams { pub gravity: BlobsVec2, pub use_spatial_hash: i32, pub device: i32, pub body_capacity_hint: u32, pub collider_capacity_hint: u32 }

#[repr(C)] #[derive(Copy, Clone, Default)]
pub struct BlobsBodyDesc {
    pub position: BlobsVec2, pub position_old: BlobsVec2, pub gravity_mod: c_float, pub rotation: c_float, pub scale: BlobsVec2,
    pub acceleration: BlobsVec2, pub velocity_request: BlobsVec2, pub calculated_velocity: BlobsVec2,
    pub has_velocity_request: i32, pub body_type: u32, pub user_data_lo: u64, pub user_data_hi: u64,
}
#[repr(C)] #[derive(Copy, Clone, Default)]
pub struct BlobsBodyState {
    pub position: BlobsVec2, pub position_old: BlobsVec2, pub center_of_mass: BlobsVec2, pub scale: BlobsVec2, pub acceleration: BlobsVec2,
    pub velocity_request: BlobsVec2, pub calculated_velocity: BlobsVec2, pub calculated_mass: c_float, pub gravity_mod: c_float,
    pub rotation: c_float, pub angular_velocity: c_float, pub torque: c_float, pub inertia: c_float, pub has_velocity_request: i32,
    pub body_type: u32, pub user_data_lo: u64, pub user_data_hi: u64,
}
#[repr(C)] #[derive(Copy, Clone, Default)]
pub struct BlobsColliderDesc {
    pub offset: BlobsAffine2, pub absolute_transform: BlobsAffine2, pub radius: c_float, pub mass_override: c_float, pub shape_radius: c_float,
    pub has_mass_override: i32, pub is_sensor: i32, pub memberships: u32, pub filter: u32, pub user_data_lo: u64, pub user_data_hi: u64,
}
#[repr(C)] #[derive(Copy, Clone, Default)] pub struct BlobsColliderState { pub desc: BlobsColliderDesc, pub parent: BlobsHandle }
#[repr(C)] #[derive(Copy, Clone, Default)] pub struct BlobsCollisionEvent { pub col_handle_a: u64, pub col_handle_b: u64, pub impact_vel_a: BlobsVec2, pub impact_vel_b: BlobsVec2 }
#[repr(C)] #[derive(Copy, Clone, Default)]
pub struct BlobsStepStats { pub collisions: u64, pub coincident_pairs: u64, pub events_dropped: u64, pub nan_detected: u32, pub steps_run: u32, pub substeps_run: u32, pub list_overflow: u32, pub gpu_ms: c_float }

pub const BLOBS_OK: i32 = 0;
pub const BLOBS_PARAM_GRAVITY_X: i32 = 0; pub const BLOBS_PARAM_GRAVITY_Y: i32 = 1; pub const BLOBS_PARAM_SUBSTEPS: i32 = 2;
pub const BLOBS_PARAM_JOINT_ITERATIONS: i32 = 3; pub const BLOBS_PARAM_USE_SPATIAL_HASH: i32 = 4; pub const BLOBS_PARAM_COLLISIONS_ENABLED: i32 = 5;
pub const BLOBS_PARAM_ACCUMULATOR: i32 = 6; pub const BLOBS_PARAM_TIME: i32 = 7; pub const BLOBS_PARAM_OLD_DT: i32 = 8; pub const BLOBS_PARAM_CELL_SIZE: i32 = 9;
pub const BLOBS_RECORD_EVENTS: i32 = 2;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct BlobsQueryFilter { pub flags: u32, pub has_groups: i32, pub memberships: u32, pub filter: u32, pub exclude_collider: BlobsHandle,
                              pub exclude_rigid_body: BlobsHandle, pub batch_world: u32, pub reserved: u32 }
pub const BLOBS_ERR_CAPACITY: i32 = 9;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct BlobsDebugCounts { pub bodies: u64, pub joints: u64, pub colliders: u64, pub springs: u64 }

/// PhysicsEvent (events.rs:42-50) as the library stores it
#[repr(C)]
#[derive(Clone, Copy)]
pub struct BlobsPhysicsEvent { pub real_time: c_double, pub unpaused_time: c_double, pub position: BlobsVec2, pub has_position: i32, pub severity: i32,
                               pub col_handle: BlobsHandle, pub rbd_handle: BlobsHandle, pub message: [c_char; 64] }

// field mask of blobs_body_set (include/blobs_b200.h)
pub const BLOBS_BODY_POSITION: u32 = 1 << 0; pub const BLOBS_BODY_POSITION_OLD: u32 = 1 << 1; pub const BLOBS_BODY_ACCELERATION: u32 = 1 << 2;
pub const BLOBS_BODY_VELOCITY_REQUEST: u32 = 1 << 3; pub const BLOBS_BODY_CALC_VELOCITY: u32 = 1 << 4; pub const BLOBS_BODY_ROTATION: u32 = 1 << 5;
pub const BLOBS_BODY_ANGULAR_VELOCITY: u32 = 1 << 6; pub const BLOBS_BODY_TORQUE: u32 = 1 << 7; pub const BLOBS_BODY_MASS: u32 = 1 << 8;
pub const BLOBS_BODY_INERTIA: u32 = 1 << 9; pub const BLOBS_BODY_GRAVITY_MOD: u32 = 1 << 10; pub const BLOBS_BODY_TYPE: u32 = 1 << 11;
pub const BLOBS_BODY_USER_DATA: u32 = 1 << 12; pub const BLOBS_BODY_SCALE: u32 = 1 << 13; pub const BLOBS_BODY_CENTER_OF_MASS: u32 = 1 << 14;

#[link(name = "blobs_b200")]
extern "C" {
    pub fn blobs_world_create(params: *const BlobsParams, out: *mut *mut BlobsWorld) -> i32;
    pub fn blobs_world_destroy(w: *mut BlobsWorld) -> i32;
    pub fn blobs_world_reset(w: *mut BlobsWorld) -> i32;
    pub fn blobs_last_error(w: *const BlobsWorld) -> *const c_char;
    pub fn blobs_world_set_param(w: *mut BlobsWorld, id: i32, value: c_double) -> i32;
    pub fn blobs_world_get_param(w: *const BlobsWorld, id: i32, out: *mut c_double) -> i32;
    pub fn blobs_body_insert(w: *mut BlobsWorld, d: *const BlobsBodyDesc, out: *mut BlobsHandle) -> i32;
    pub fn blobs_body_remove(w: *mut BlobsWorld, h: BlobsHandle) -> i32;
    pub fn blobs_body_get(w: *mut BlobsWorld, h: BlobsHandle, out: *mut BlobsBodyState) -> i32;
    pub fn blobs_body_set(w: *mut BlobsWorld, h: BlobsHandle, s: *const BlobsBodyState, mask: u32) -> i32;
    pub fn blobs_body_count(w: *const BlobsWorld, out: *mut u64) -> i32;
    pub fn blobs_body_translate(w: *mut BlobsWorld, h: BlobsHandle, off: BlobsVec2) -> i32;
    pub fn blobs_collider_insert(w: *mut BlobsWorld, d: *const BlobsColliderDesc, parent: BlobsHandle, out: *mut BlobsHandle) -> i32;
    pub fn blobs_collider_remove(w: *mut BlobsWorld, h: BlobsHandle) -> i32;
    pub fn blobs_collider_get(w: *mut BlobsWorld, h: BlobsHandle, out: *mut BlobsColliderState) -> i32;
    pub fn blobs_spring_insert(w: *mut BlobsWorld, a: BlobsHandle, b: BlobsHandle, rest: c_float, k: c_float, c: c_float, out: *mut BlobsHandle) -> i32;
    pub fn blobs_spring_remove(w: *mut BlobsWorld, h: BlobsHandle) -> i32;                                            // physics.springs.remove(index) demos/joints.rs:79
    pub fn blobs_joint_insert(w: *mut BlobsWorld, a: BlobsHandle, b: BlobsHandle, aa: BlobsVec2, ab: BlobsVec2, dist_or_nan: c_float, out: *mut BlobsHandle) -> i32;
    pub fn blobs_constraint_push(w: *mut BlobsWorld, p: BlobsVec2, r: c_float) -> i32;
    pub fn blobs_constraint_clear(w: *mut BlobsWorld) -> i32;
    pub fn blobs_step(w: *mut BlobsWorld, delta: c_double, stats: *mut BlobsStepStats) -> i32;
    pub fn blobs_fixed_step(w: *mut BlobsWorld, frame_time: c_double, stats: *mut BlobsStepStats) -> i32;
    pub fn blobs_record_contacts(w: *mut BlobsWorld, mode: c_int, capacity: usize) -> i32;
    pub fn blobs_events_drain(w: *mut BlobsWorld, buf: *mut BlobsCollisionEvent, cap: usize, n: *mut usize) -> i32;
    pub fn blobs_download_bodies(w: *mut BlobsWorld, states: *mut BlobsBodyState, handles: *mut BlobsHandle, cap: usize) -> i32;
    pub fn blobs_download_colliders(w: *mut BlobsWorld, states: *mut BlobsColliderState, handles: *mut BlobsHandle, cap: usize) -> i32;
    pub fn blobs_query_circles(w: *mut BlobsWorld, n: usize, centre_xy: *const c_float, radius: *const c_float, filter: *const BlobsQueryFilter,
                               offsets: *mut u64, hits: *mut BlobsHandle, hit_cap: usize, n_hits: *mut usize) -> i32;
    pub fn blobs_debug_counts(w: *const BlobsWorld, out: *mut BlobsDebugCounts) -> i32;
    pub fn blobs_debug_data(w: *mut BlobsWorld, body_xform: *mut c_float, joint_ab: *mut c_float, col_xform: *mut c_float, col_radius: *mut c_float,
                            spring_ab: *mut c_float, caps: *const BlobsDebugCounts) -> i32;
    pub fn blobs_body_slots(w: *const BlobsWorld, out: *mut u64) -> i32;
    pub fn blobs_collider_slots(w: *const BlobsWorld, out: *mut u64) -> i32;
    // perf_counters.rs:52-87 (process-global registry; blobs_step* feeds "collisions")
    pub fn blobs_perf_counter(name: *const c_char, count: u64);
    pub fn blobs_perf_counter_inc(name: *const c_char, inc: u64);
    pub fn blobs_perf_counters_new_frame(delta: c_double);
    pub fn blobs_perf_counters_reset();
    pub fn blobs_perf_counter_get(name: *const c_char, count: *mut u64, decayed_average: *mut c_double) -> i32;
    pub fn blobs_perf_counter_count() -> u64;
    pub fn blobs_perf_counter_at(i: u64, name: *mut c_char, name_cap: usize, count: *mut u64, decayed_average: *mut c_double) -> i32;
    // events.rs:20-64 (process-global soft-error ring)
    pub fn blobs_event_history_len() -> u64;
    pub fn blobs_event_history_get(i: u64, out: *mut BlobsPhysicsEvent) -> i32;
    pub fn blobs_event_history_clear();
}
