/* blobs_b200.h — C ABI of libblobs_b200.so: the B200-native drop-in for the per-step hot path of
 * darthdeus/blobs (`blobs::Physics::step`, reference file blobs/src/physics.rs).
 *
 * The reference has no FFI layer; its boundary is the public Rust surface of `Physics`
 * (physics.rs:36-239), the builders (rigid_body.rs:302-401, collider.rs:211-284) and the handle
 * types. Each entry point below names the reference item it replaces (file:line relative to the
 * reference checkout). A thin Rust shim (rust/) maps these 1:1 back onto the
 * reference's names so `Physics::step` stays a drop-in.
 *
 * Conventions: plain pointers and sizes only; every call returns a BlobsStatus (0 = OK) unless noted;
 * descriptors are copied, no borrowed pointer survives a call; a world is owned by one host thread
 * (the reference's Physics is !Send + !Sync, physics.rs:4). There is NO CPU fallback: if no CUDA
 * device is usable, blobs_world_create fails with BLOBS_ERR_CUDA.
 *
 * Handles are thunderdome::Index::to_bits(): generation << 32 | slot; 0 is never a valid handle.
 */
#ifndef BLOBS_B200_H
#define BLOBS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLOBS_ABI_VERSION 1

typedef struct BlobsWorld BlobsWorld;
typedef uint64_t BlobsHandle;

typedef enum BlobsStatus {
    BLOBS_OK = 0,
    BLOBS_ERR_STALE_HANDLE = 1,   /* Option::None in the reference (arena lookup failed) */
    BLOBS_ERR_SAME_BODY = 2,      /* thunderdome get2_mut panic: identical indices (physics.rs:191-196, springs.rs:26-29) */
    BLOBS_ERR_NAN = 3,            /* assert!(!n.is_nan()) physics.rs:293; rotation asserts physics.rs:471-474 */
    BLOBS_ERR_CUDA = 4,           /* CUDA runtime failure; see blobs_last_error */
    BLOBS_ERR_INVALID = 5,        /* bad argument */
    BLOBS_ERR_SPATIAL_HASH = 6,   /* panic!("spatial collisions not supported right now") physics.rs:412 */
    BLOBS_ERR_MASS = 7,           /* assert!(calculated_mass > 0.0) physics.rs:447-448 */
    BLOBS_ERR_DANGLING = 8,       /* spring/joint references a removed body: unwrap() panic physics.rs:427-432, springs.rs:26-29 */
    BLOBS_ERR_CAPACITY = 9
} BlobsStatus;

typedef struct BlobsVec2 { float x, y; } BlobsVec2;

/* glam::Affine2, column-major like glam: matrix2.x_axis, matrix2.y_axis, translation */
typedef struct BlobsAffine2 { BlobsVec2 x_axis, y_axis, translation; } BlobsAffine2;

/* RigidBodyType, rigid_body.rs:221-242 */
enum { BLOBS_BODY_DYNAMIC = 0, BLOBS_BODY_STATIC = 1, BLOBS_BODY_KINEMATIC_POSITION = 2, BLOBS_BODY_KINEMATIC_VELOCITY = 3 };

/* Physics::new(gravity, use_spatial_hash), physics.rs:37-69; defaults physics.rs:46-47,62-67 */
typedef struct BlobsParams {
    BlobsVec2 gravity;
    int32_t use_spatial_hash;    /* kept for API parity; stepping with it set returns BLOBS_ERR_SPATIAL_HASH */
    int32_t device;              /* CUDA device ordinal, -1 = current */
    uint32_t body_capacity_hint; /* 0 = grow on demand */
    uint32_t collider_capacity_hint;
} BlobsParams;

/* Runtime knobs that are `pub` fields on Physics (physics.rs:6-33) */
typedef enum BlobsParamId {
    BLOBS_PARAM_GRAVITY_X = 0,
    BLOBS_PARAM_GRAVITY_Y = 1,
    BLOBS_PARAM_SUBSTEPS = 2,            /* physics.rs:8  (default 8) */
    BLOBS_PARAM_JOINT_ITERATIONS = 3,    /* physics.rs:9  (default 4) */
    BLOBS_PARAM_USE_SPATIAL_HASH = 4,    /* physics.rs:20 */
    BLOBS_PARAM_COLLISIONS_ENABLED = 5,  /* physics.rs:27 */
    BLOBS_PARAM_ACCUMULATOR = 6,         /* physics.rs:30 */
    BLOBS_PARAM_TIME = 7,                /* physics.rs:31 */
    BLOBS_PARAM_OLD_DT = 8,              /* physics.rs:33 */
    BLOBS_PARAM_CELL_SIZE = 9,           /* spatial_hash.cell_size, spatial.rs:32 (default 2.0, physics.rs:66) */
    BLOBS_PARAM_BROADPHASE_CELL = 10,    /* GPU grid cell edge; 0 = auto (2 * max collider radius) */
    BLOBS_PARAM_CONTACT_MODE = 11,       /* 0 = ordered (bit-exact summation order), 1 = fast (unordered) */
    BLOBS_PARAM_FUSED = 12,              /* 1 = allow the fused contact+verlet kernel (default), 0 = force split kernels */
    BLOBS_PARAM_TUNE = 13,               /* kernel-variant selector for benchmarking (0 = default); never changes results. List pipeline:
                                            CTAs of k_step per SM - 1: 3, 2: 5, 3: 6, 4: 7, 5: 8, 6: 4 (default 5; 6 on strips) */
    BLOBS_PARAM_BATCH_WORLD = 14         /* batched independent worlds (BASELINE config #3): bodies inserted from now on belong to this
                                            world id; worlds never interact, each one behaves like its own Physics (gravity, constraints,
                                            substeps are shared). Default 0 = the single world. */
    ,BLOBS_PARAM_GRAPH = 15              /* 1 (default): a whole Physics::integrate call is captured once as a CUDA graph and replayed while
                                            nothing structural changes; 0: plain launches. Never changes results. */
    ,BLOBS_PARAM_GRAPH_REPLAYS = 16      /* read-only: number of graph replays so far */
    ,BLOBS_PARAM_STRIP_MAX_GHOSTS = 17   /* read-only: largest ghost / migrant section received from a neighbour so far (strip mode); */
    ,BLOBS_PARAM_STRIP_MAX_MIGRANTS = 18 /*            size the capacities of blobs_strip_configure from these */
    ,BLOBS_PARAM_CROWDED = 19            /* bodies with more contacts than the in-register ordered list holds (24): 0 = resolved inline by
                                            their own thread, 1 = deferred to a warp-per-body kernel, 2 (default) = automatic (deferred
                                            once a step has seen such bodies). Never changes results. */
    ,BLOBS_PARAM_POOL = 20               /* contact resolution inside k_main: 0 = per lane, 1 = pooled across the warp (one body-candidate
                                            pair per lane, rank-ordered sums; for contact-rich states), 2 (default) = automatic, by the
                                            contact density of the previous step call. Never changes results. */
    ,BLOBS_PARAM_POOL_MIN = 21           /* pooled path: minimum prefilter survivors in a warp (default 16) */
    ,BLOBS_PARAM_STRIP_P2P = 22          /* strip mode, set BEFORE blobs_strip_configure: 1 = exchange ghosts / migrants by direct peer-memory
                                            stores over NVLink (neighbours' receive buffers mapped through CUDA IPC, one push kernel per
                                            substep) instead of grouped ncclSend/ncclRecv; falls back to NCCL on every rank if any rank
                                            cannot map its neighbours. Reads back 1 only while the peer path is active. Never changes results. */
    ,BLOBS_PARAM_LIST = 23               /* broadphase strategy: 1 = per-collider neighbour lists (every collider within r_a + r_b + skin, sorted by
                                            slot), walked by one fused kernel per substep and rebuilt from the cell grid only when the device-side
                                            displacement tracking says a pair outside the lists could touch; 0 = the cell grid is rebuilt every
                                            substep; 2 (default) = lists, falling back to 0 for a while whenever the scene is so agitated that the
                                            lists are rebuilt almost every substep. The contact set of every substep is the reference's either way
                                            (physics.rs:241-317): never changes results. */
    ,BLOBS_PARAM_SKIN = 24               /* neighbour-list skin as a fraction of the largest collider radius (default 0.8) */
    ,BLOBS_PARAM_LIST_REBUILDS = 25      /* read-only: list rebuilds / substeps run so far (as of the last blobs_step* call) */
    ,BLOBS_PARAM_LIST_SUBSTEPS = 26
    ,BLOBS_PARAM_LIST_ACTIVE = 27        /* read-only: 1 while the neighbour-list pipeline is the one in use */
} BlobsParamId;

/* RigidBodyBuilder, rigid_body.rs:287-401 */
typedef struct BlobsBodyDesc {
    BlobsVec2 position;
    BlobsVec2 position_old;
    float gravity_mod;
    float rotation;
    BlobsVec2 scale;
    BlobsVec2 acceleration;
    BlobsVec2 velocity_request;
    BlobsVec2 calculated_velocity;
    int32_t has_velocity_request;
    uint32_t body_type;
    uint64_t user_data_lo, user_data_hi;
} BlobsBodyDesc;

/* RigidBody, rigid_body.rs:41-74 (all fields are pub in the reference) */
typedef struct BlobsBodyState {
    BlobsVec2 position;
    BlobsVec2 position_old;
    BlobsVec2 center_of_mass;
    BlobsVec2 scale;
    BlobsVec2 acceleration;
    BlobsVec2 velocity_request;
    BlobsVec2 calculated_velocity;
    float calculated_mass;
    float gravity_mod;
    float rotation;
    float angular_velocity;
    float torque;
    float inertia;
    int32_t has_velocity_request;
    uint32_t body_type;
    uint64_t user_data_lo, user_data_hi;
} BlobsBodyState;

/* field mask for blobs_body_set (the reference mutates through get_mut_rbd, physics.rs:109-111) */
enum {
    BLOBS_BODY_POSITION = 1u << 0,
    BLOBS_BODY_POSITION_OLD = 1u << 1,
    BLOBS_BODY_ACCELERATION = 1u << 2,
    BLOBS_BODY_VELOCITY_REQUEST = 1u << 3,  /* set_velocity, rigid_body.rs:182-184 */
    BLOBS_BODY_CALC_VELOCITY = 1u << 4,
    BLOBS_BODY_ROTATION = 1u << 5,
    BLOBS_BODY_ANGULAR_VELOCITY = 1u << 6,
    BLOBS_BODY_TORQUE = 1u << 7,
    BLOBS_BODY_MASS = 1u << 8,
    BLOBS_BODY_INERTIA = 1u << 9,
    BLOBS_BODY_GRAVITY_MOD = 1u << 10,
    BLOBS_BODY_TYPE = 1u << 11,
    BLOBS_BODY_USER_DATA = 1u << 12,
    BLOBS_BODY_SCALE = 1u << 13,
    BLOBS_BODY_CENTER_OF_MASS = 1u << 14,
    BLOBS_BODY_ALL = 0x7fffu
};

/* ColliderBuilder, collider.rs:199-284. Shapes: only Ball exists in the reference (lib.rs:49-74). */
typedef struct BlobsColliderDesc {
    BlobsAffine2 offset;
    BlobsAffine2 absolute_transform;
    float radius;
    float mass_override;
    float shape_radius;          /* Ball::radius of `shape` (debug/AABB only, lib.rs:61-67) */
    int32_t has_mass_override;
    int32_t is_sensor;
    uint32_t memberships, filter; /* InteractionGroups, groups.rs:7-12 */
    uint64_t user_data_lo, user_data_hi;
} BlobsColliderDesc;

typedef struct BlobsColliderState {
    BlobsColliderDesc desc;      /* absolute_transform is the live snapshot (physics.rs:360-366) */
    BlobsHandle parent;          /* 0 = None */
} BlobsColliderState;

/* CollisionEvent, lib.rs:146-153 (sent once per contact per substep, physics.rs:304-311) */
typedef struct BlobsCollisionEvent {
    BlobsHandle col_handle_a, col_handle_b;
    BlobsVec2 impact_vel_a, impact_vel_b;
} BlobsCollisionEvent;

typedef struct BlobsStepStats {
    uint64_t collisions;        /* perf_counter_inc("collisions", count), physics.rs:316 — summed over substeps */
    uint64_t coincident_pairs;  /* pairs that took the distance < 1e-6 branch, physics.rs:272-286 */
    uint64_t events_dropped;    /* events/pairs that did not fit the recording buffer */
    uint32_t nan_detected;      /* bit 0: a position became NaN; bit 1: a rotation became non-finite (the reference would have panicked,
                                   physics.rs:471-474); strip mode - bit 2: a ghost / migration message or the owned list overflowed its
                                   capacity, bit 3: a peer-memory exchange waited > ~4 s for a neighbour (results are not valid) */
    uint32_t steps_run;         /* integrate() calls performed (fixed_step: 0..3, physics.rs:88-98) */
    uint32_t substeps_run;
    uint32_t list_overflow;     /* contact lists that exceeded the in-register capacity and took the rescan path */
    float gpu_ms;               /* CUDA-event time of the kernels of this call */
} BlobsStepStats;

/* ---- world lifetime ------------------------------------------------------------------------ */
int32_t blobs_abi_version(void);
int32_t blobs_world_create(const BlobsParams* params, BlobsWorld** out);      /* Physics::new physics.rs:37 */
int32_t blobs_world_destroy(BlobsWorld* w);
int32_t blobs_world_reset(BlobsWorld* w);                                     /* Physics::reset physics.rs:71-76 */
const char* blobs_last_error(const BlobsWorld* w);                            /* panic/expect message equivalent */
int32_t blobs_world_set_param(BlobsWorld* w, int32_t id, double value);       /* pub fields physics.rs:6-33 */
int32_t blobs_world_get_param(const BlobsWorld* w, int32_t id, double* out);

/* ---- bodies ---------------------------------------------------------------------------------- */
int32_t blobs_body_insert(BlobsWorld* w, const BlobsBodyDesc* desc, BlobsHandle* out);          /* insert_rbd physics.rs:121-128 */
int32_t blobs_body_insert_many(BlobsWorld* w, size_t n, const BlobsBodyDesc* descs, BlobsHandle* out); /* n x insert_rbd */
int32_t blobs_body_remove(BlobsWorld* w, BlobsHandle h);                                       /* remove_rbd physics.rs:163-172 */
int32_t blobs_body_get(BlobsWorld* w, BlobsHandle h, BlobsBodyState* out);                      /* get_rbd physics.rs:105-107 */
int32_t blobs_body_set(BlobsWorld* w, BlobsHandle h, const BlobsBodyState* s, uint32_t mask);   /* get_mut_rbd physics.rs:109-111 */
int32_t blobs_body_count(const BlobsWorld* w, uint64_t* out);                                   /* rbd_count physics.rs:113-115 */
int32_t blobs_body_translate(BlobsWorld* w, BlobsHandle h, BlobsVec2 offset);                   /* update_rigid_body_position physics.rs:174-182 */
int32_t blobs_body_apply_force(BlobsWorld* w, BlobsHandle h, BlobsVec2 force);                  /* RigidBody::apply_force rigid_body.rs:155-160 */
int32_t blobs_body_colliders(const BlobsWorld* w, BlobsHandle h, BlobsHandle* out, size_t cap, size_t* n); /* rbd.colliders (each handle appears twice) */

/* ---- colliders --------------------------------------------------------------------------------- */
int32_t blobs_collider_insert(BlobsWorld* w, const BlobsColliderDesc* desc, BlobsHandle parent, BlobsHandle* out); /* insert_collider_with_parent physics.rs:130-149 */
int32_t blobs_collider_insert_many(BlobsWorld* w, size_t n, const BlobsColliderDesc* descs, const BlobsHandle* parents, BlobsHandle* out);
int32_t blobs_collider_remove(BlobsWorld* w, BlobsHandle h);                                   /* remove_col physics.rs:159-161 */
int32_t blobs_collider_get(BlobsWorld* w, BlobsHandle h, BlobsColliderState* out);              /* get_col physics.rs:117-119 */
int32_t blobs_collider_count(const BlobsWorld* w, uint64_t* out);

/* ---- springs, joints, constraints ------------------------------------------------------------- */
int32_t blobs_spring_insert(BlobsWorld* w, BlobsHandle a, BlobsHandle b, float rest_length, float stiffness, float damping, BlobsHandle* out); /* physics.springs.insert(Spring{..}) springs.rs:16-22 */
int32_t blobs_spring_remove(BlobsWorld* w, BlobsHandle h);
/* create_fixed_joint (physics.rs:184-207) when distance is NaN, else create_fixed_joint_with_distance (physics.rs:209-239) */
int32_t blobs_joint_insert(BlobsWorld* w, BlobsHandle a, BlobsHandle b, BlobsVec2 anchor_a, BlobsVec2 anchor_b, float distance_or_nan, BlobsHandle* out);
int32_t blobs_joint_remove(BlobsWorld* w, BlobsHandle h);
/* bulk forms: n x springs.insert / n x create_fixed_joint(_with_distance); params = (rest_length, stiffness, damping) per spring,
 * anchors = (anchor_a.xy, anchor_b.xy) per joint or NULL for zeros, distance_or_nan NULL = all NaN (distance from positions) */
int32_t blobs_spring_insert_many(BlobsWorld* w, size_t n, const BlobsHandle* a, const BlobsHandle* b, const float* params3, BlobsHandle* out);
int32_t blobs_joint_insert_many(BlobsWorld* w, size_t n, const BlobsHandle* a, const BlobsHandle* b, const float* anchors4, const float* distance_or_nan, BlobsHandle* out);
int32_t blobs_constraint_push(BlobsWorld* w, BlobsVec2 position, float radius);                 /* physics.constraints.push(Constraint{..}) lib.rs:189-193 */
int32_t blobs_constraint_clear(BlobsWorld* w);

/* ---- stepping ----------------------------------------------------------------------------------- */
int32_t blobs_step(BlobsWorld* w, double delta, BlobsStepStats* stats);             /* Physics::step physics.rs:78-82 */
int32_t blobs_fixed_step(BlobsWorld* w, double frame_time, BlobsStepStats* stats);  /* Physics::fixed_step physics.rs:84-99 */
/* n back-to-back step(delta) calls enqueued without host synchronisation in between (throughput path) */
int32_t blobs_step_n(BlobsWorld* w, double delta, uint32_t n, BlobsStepStats* stats);

/* ---- bulk state transfer (slot-indexed; slot = low 32 bits of the handle) --------------------- */
int32_t blobs_body_slots(const BlobsWorld* w, uint64_t* out);      /* arena storage length, incl. free slots */
int32_t blobs_collider_slots(const BlobsWorld* w, uint64_t* out);
/* handles[s] = 0 for a free slot; either pointer may be NULL */
int32_t blobs_download_bodies(BlobsWorld* w, BlobsBodyState* states, BlobsHandle* handles, size_t cap);
int32_t blobs_download_colliders(BlobsWorld* w, BlobsColliderState* states, BlobsHandle* handles, size_t cap);
/* raw SoA fast paths: xy interleaved, `cap` slots; host buffers may be pinned */
int32_t blobs_read_body_positions(BlobsWorld* w, float* xy, size_t cap);
int32_t blobs_read_body_velocities(BlobsWorld* w, float* xy, size_t cap);
int32_t blobs_apply_forces(BlobsWorld* w, const float* force_xy, size_t cap);  /* per-slot RigidBody::apply_force rigid_body.rs:155-160 */
/* Pipelined forms of the two calls above for a per-frame loop (no reference counterpart: the reference's state is host memory).
 * The PCIe copies run on their own streams, one per direction, so they overlap the kernels of the neighbouring steps:
 *     blobs_forces_upload_async(w, f[0]);
 *     for each frame i:  blobs_apply_forces_uploaded(w);             // forces i (already on the device)
 *                        blobs_forces_upload_async(w, f[i+1]);        // travels while step i computes
 *                        blobs_step(w, delta, &stats);
 *                        blobs_io_sync(w);                            // positions i-1 have arrived, f[i+1] may be reused
 *                        blobs_read_body_positions_async(w, out[i&1]); // travels while step i+1 computes
 *     blobs_io_sync(w);
 * Host buffers should be pinned (cudaHostAlloc / torch pin_memory) and must not be touched between the call and the next
 * blobs_io_sync. Results are identical to the synchronous calls. */
int32_t blobs_forces_upload_async(BlobsWorld* w, const float* force_xy, size_t cap);  /* at most one batch may be pending */
int32_t blobs_apply_forces_uploaded(BlobsWorld* w);                                   /* per-slot RigidBody::apply_force rigid_body.rs:155-160 */
int32_t blobs_read_body_positions_async(BlobsWorld* w, float* xy, size_t cap);
int32_t blobs_io_sync(BlobsWorld* w);
/* SpatialHash::get_cell_coords (spatial.rs:57-62) of every collider snapshot, with BLOBS_PARAM_CELL_SIZE */
int32_t blobs_download_cell_coords(BlobsWorld* w, int32_t* cx, int32_t* cy, size_t cap);

/* Physics::debug_data (physics.rs:479-481) / make_debug_data (debug.rs:34-91): one call, arena iteration order (ascending
 * slot, free slots skipped). body_xform / col_xform: 6 floats per entry = glam::Affine2 (x_axis.xy, y_axis.xy, translation.xy),
 * bodies as from_angle_translation(rotation, position), colliders as collider.absolute_transform; col_radius: Ball radius;
 * joint_ab / spring_ab: 4 floats per entry = position of rigid_body_a, position of rigid_body_b (NaN where the reference would
 * panic on a removed body). Any pointer may be NULL; `caps` holds the capacities in entries. */
typedef struct BlobsDebugCounts { uint64_t bodies, joints, colliders, springs; } BlobsDebugCounts;
int32_t blobs_debug_counts(const BlobsWorld* w, BlobsDebugCounts* out);
int32_t blobs_debug_data(BlobsWorld* w, float* body_xform, float* joint_ab, float* col_xform, float* col_radius, float* spring_ab,
                         const BlobsDebugCounts* caps);

/* ---- scene queries: the reference's SpatialHash::query (spatial.rs:155-195) and the QueryPipeline / QueryFilter it
 * left as stubs (lib.rs:167-187, query_filter.rs:27-108), served from the GPU broadphase table. n circle queries are answered
 * in one call against the live collider snapshots: collider c is a hit iff |c.position - centre|^2 <= (radius + c.radius)^2
 * (inclusive, as SpatialHash::query) and it passes the filter. Colliders whose parent body is gone are never reported.
 * hits of query q = hits[offsets[q] .. offsets[q+1]), ascending slot order. If hit_cap is too small nothing is written,
 * *n_hits holds the required capacity and BLOBS_ERR_CAPACITY is returned. Not available in strip mode. */
enum { BLOBS_QUERY_EXCLUDE_FIXED = 1u << 1, BLOBS_QUERY_EXCLUDE_KINEMATIC = 1u << 2, BLOBS_QUERY_EXCLUDE_DYNAMIC = 1u << 3,
       BLOBS_QUERY_EXCLUDE_SENSORS = 1u << 4, BLOBS_QUERY_EXCLUDE_SOLIDS = 1u << 5 };   /* QueryFilterFlags, query_filter.rs:6-25 */
typedef struct BlobsQueryFilter {
    uint32_t flags;                      /* BLOBS_QUERY_* */
    int32_t has_groups;                  /* QueryFilter::groups is Some */
    uint32_t memberships, filter;        /* ... tested with InteractionGroups::test against the collider's groups */
    BlobsHandle exclude_collider;        /* 0 = None */
    BlobsHandle exclude_rigid_body;      /* 0 = None */
    uint32_t batch_world;                /* which batched world to query (BLOBS_PARAM_BATCH_WORLD), 0 = the single world */
    uint32_t reserved;
} BlobsQueryFilter;
int32_t blobs_query_circles(BlobsWorld* w, size_t n, const float* centre_xy, const float* radius, const BlobsQueryFilter* filter_or_null,
                            uint64_t* offsets /* n + 1 */, BlobsHandle* hits, size_t hit_cap, size_t* n_hits);

/* ---- contact output ------------------------------------------------------------------------------ */
enum { BLOBS_RECORD_OFF = 0, BLOBS_RECORD_PAIRS = 1, BLOBS_RECORD_EVENTS = 2 };
/* collision_send / collision_recv, physics.rs:22-23,304-311. PAIRS records slots only. */
int32_t blobs_record_contacts(BlobsWorld* w, int32_t mode, size_t capacity);
/* *n = events pending before the call; the first min(*n, cap) are returned and consumed, the rest stay queued for the next call
 * (loop while *n > cap). Events beyond the recording capacity are counted in BlobsStepStats::events_dropped, never silently lost. */
int32_t blobs_events_drain(BlobsWorld* w, BlobsCollisionEvent* buf, size_t cap, size_t* n);
/* slot pairs (a > b) recorded since the last drain; substep_end[i] = running pair count after substep i */
int32_t blobs_pairs_drain(BlobsWorld* w, uint32_t* slot_a, uint32_t* slot_b, size_t cap, size_t* n,
                          uint64_t* substep_end, size_t substep_cap, size_t* n_substeps);

/* ---- introspection used by bench.py / tests ---------------------------------------------------- */
typedef struct BlobsKernelInfo {
    uint64_t launches;          /* kernels of this library launched so far on this world */
    uint32_t grid_w, grid_h;    /* broadphase table dims */
    float broadphase_cell;
    float r_max;
    uint32_t fused_path;        /* 1 if the last step used the fused contact+verlet kernel */
    uint32_t n_simple_bodies, n_multi_bodies, n_spring_bodies, n_islands;
} BlobsKernelInfo;
int32_t blobs_kernel_info(const BlobsWorld* w, BlobsKernelInfo* out);
/* CUDA-event timing of individual kernel classes during the next steps (0 = off, 1 = every kernel class, 2 = the dominant kernel
 * (contact + update kernel and k_crowded) only, so that the rest of the step runs unperturbed). Used for the roofline. */
int32_t blobs_profile_enable(BlobsWorld* w, int32_t on);
/* ms accumulated per kernel class since enable: [0]=main/contacts, [1]=scan, [2]=scatter, [3]=springs, [4]=joints, [5]=integrate, [6]=other,
 * [7]=strip pack, [8]=strip ghost binning/scatter/hand-over, [9]=ghost exchange, [10]=k_crowded, [11]=neighbour-list build, [12]=list decision */
int32_t blobs_profile_read(BlobsWorld* w, float* ms, uint64_t* launches, size_t n);

/* ---- perf counters: the reference's process-global registry (perf_counters.rs:3-87). blobs_step* feeds "collisions"
 * (physics.rs:316); the host application calls new_frame(delta) once per frame (demo/src/main.rs:223) and reads
 * (count, decayed_average) pairs for its perf panel (main.rs:291-300). Not per world: one registry per process, like the
 * reference's static. Kernels are also bracketed by profiler (NVTX) ranges named after the reference's tracy spans
 * ("step", "integrate", "substep", "brute_force_collisions", "update positions"; physics.rs:79,92,242,324,398,402). */
void blobs_perf_counter(const char* name, uint64_t count);                      /* perf_counter, perf_counters.rs:66-69 */
void blobs_perf_counter_inc(const char* name, uint64_t inc);                    /* perf_counter_inc, perf_counters.rs:71-76 */
void blobs_perf_counters_new_frame(double delta);                               /* perf_counters_new_frame, perf_counters.rs:56-59 */
void blobs_perf_counters_reset(void);                                           /* reset_perf_counters, perf_counters.rs:61-64 */
int32_t blobs_perf_counter_get(const char* name, uint64_t* count, double* decayed_average); /* get_perf_counter, perf_counters.rs:78-81: (0, 0.0) if absent */
uint64_t blobs_perf_counter_count(void);                                        /* PerfCounters::global().counters.len() */
/* i-th counter in name order (the reference iterates a HashMap: unordered); BLOBS_ERR_INVALID past the end, BLOBS_ERR_CAPACITY if the
 * name (with its terminating NUL) does not fit name_cap */
int32_t blobs_perf_counter_at(uint64_t i, char* name, size_t name_cap, uint64_t* count, double* decayed_average);

/* ---- soft-error history: the reference's process-global event ring (events.rs:20-64; at most 1000 entries, oldest dropped).
 * Two messages exist: "removing a non-existent rigid body" (Error; remove_rbd on a stale handle, rigid_body.rs:266-275 - the call
 * itself returns BLOBS_ERR_STALE_HANDLE, which the shim ignores like the reference) and "rbd removed because colliders.len() == 0"
 * (Info; removing a body's last collider removes the body, collider.rs:143-158). The reference never reads the ring back; the
 * accessors below are what a debug panel would need. */
enum { BLOBS_SEVERITY_TRACE = 0, BLOBS_SEVERITY_DEBUG, BLOBS_SEVERITY_INFO, BLOBS_SEVERITY_WARN, BLOBS_SEVERITY_ERROR, BLOBS_SEVERITY_CRITICAL }; /* events.rs:52-60 */
typedef struct BlobsPhysicsEvent {       /* PhysicsEvent, events.rs:42-50 */
    double real_time, unpaused_time;     /* TimeData (never advanced by the reference: always 0) */
    BlobsVec2 position;
    int32_t has_position;
    int32_t severity;
    BlobsHandle col_handle, rbd_handle;  /* 0 = None */
    char message[64];
} BlobsPhysicsEvent;
uint64_t blobs_event_history_len(void);
int32_t blobs_event_history_get(uint64_t i, BlobsPhysicsEvent* out);   /* i = 0 is the oldest entry still held */
void blobs_event_history_clear(void);

/* ---- multi-GPU: one large world split into vertical strips, one rank (process + GPU) per strip (BASELINE config #5).
 * No reference counterpart (the reference is single-threaded). Every rank builds the SAME full scene (identical handles),
 * then calls blobs_strip_configure; from then on blobs_step* is collective: each rank advances the bodies whose collider
 * snapshot x lies in [x_lo, x_hi) and exchanges ghost records / migrating bodies with its two neighbours once per
 * substep (grouped ncclSend/ncclRecv over NVLink). Results are bit-identical to the single-GPU world. */
int32_t blobs_strip_unique_id(uint8_t out128[128]);   /* ncclGetUniqueId on one rank; broadcast it (e.g. torch.distributed) */
int32_t blobs_strip_configure(BlobsWorld* w, int32_t rank, int32_t nranks, float x_lo, float x_hi, const uint8_t id128[128],
                              uint32_t ghost_capacity, uint32_t migrate_capacity);
int32_t blobs_strip_owned(BlobsWorld* w, uint8_t* owned_by_body_slot, size_t cap);   /* 1 = this rank currently owns the body */
/* distributed host I/O: compact (slot, position) list of the bodies this rank owns (all live bodies without strips), and the
 * matching indexed RigidBody::apply_force (rigid_body.rs:155-160; entries for bodies owned elsewhere are ignored) */
int32_t blobs_read_owned_positions(BlobsWorld* w, uint32_t* slots, float* xy, size_t cap, size_t* n);
int32_t blobs_apply_forces_indexed(BlobsWorld* w, const uint32_t* slots, const float* force_xy, size_t n);
/* Pipelined forms of the two calls above, for the per-frame loop of a strip-decomposed world (same copy streams and rules as
 * blobs_forces_upload_async / blobs_read_body_positions_async; blobs_io_sync completes them):
 *     blobs_forces_indexed_upload_async(w, slots, f, n);
 *     for each frame i:  blobs_apply_forces_indexed_uploaded(w);
 *                        blobs_forces_indexed_upload_async(w, slots', f', n');      // travels while step i computes
 *                        blobs_step(w, delta, &stats);
 *                        blobs_io_sync(w);                                          // (slots, xy, n) of frame i-1 have arrived
 *                        blobs_read_owned_positions_async(w, slots_out[i&1], xy_out[i&1], &n_out[i&1], cap);
 *     blobs_io_sync(w);
 * `n_out` is written by the device-to-host copy (keep it in pinned memory); the number of entries copied is min(cap, this rank's
 * current owned-list bound), of which the first *n_out are valid. */
int32_t blobs_forces_indexed_upload_async(BlobsWorld* w, const uint32_t* slots, const float* force_xy, size_t n);  /* at most one batch may be pending */
int32_t blobs_apply_forces_indexed_uploaded(BlobsWorld* w);
int32_t blobs_read_owned_positions_async(BlobsWorld* w, uint32_t* slots, float* xy, uint32_t* n_out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* BLOBS_B200_H */
