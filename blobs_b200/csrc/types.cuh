// POD types shared by the host orchestration (world.cu, capi.cu) and the device code (kernels.cuh).
// No functions here, so it can be included from several translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace blobs {

// body flags (host-authoritative, bflags[])
enum : uint32_t {
    BF_ALIVE = 1u << 0,
    BF_STATIC = 1u << 1,       // RigidBodyType::Static
    BF_SPRINGS = 1u << 2,      // has incident springs: gravity + spring forces are summed by k_springs
    BF_ROT = 1u << 3,          // angular state may be non-zero / rotation matters (torque, joints, rotated)
    BF_JOINTED = 1u << 4,
    BF_FIRST_DYN = 1u << 5,    // first non-static body of its world in arena order: sees dt/old_dt (physics.rs:338-339, Q2)
    BF_KINEMATIC = 1u << 6,    // RigidBodyType::KinematicPositionBased / KinematicVelocityBased (only scene queries look at it)
    BF_LOOSE = 1u << 7,        // no collider and no joint: no record stands for it in the cell-sorted array (k_tile advances these with a k_integrate pass)
};
// collider flags (host-authoritative, cflags[])
enum : uint32_t {
    CF_ACTIVE = 1u << 0,  // alive AND parent handle resolves to a live body (physics.rs:252-253,270)
    CF_SENSOR = 1u << 1,
    CF_OFFSET = 1u << 2,  // offset.translation != 0 (snapshot needs coff)
    CF_COLD = 1u << 3,    // the record's cold half is NOT the default (m = 4r, groups = ALL, sole collider of its parent)
};
constexpr uint32_t NO_SLOT = 0xffffffffu;
constexpr uint32_t OVER_MULTI_BIT = 0x80000000u;   // k_crowded list entry: index into the multi-collider body list
constexpr uint32_t OVER_COUNT_BIT = 0x40000000u;   // k_crowded list entry: pairs of this body were not counted / recorded yet
// hot.w packing: collider slot | needs-cold << 30 | is_sensor << 31
constexpr uint32_t HOT_SLOT_MASK = 0x3fffffffu, HOT_COLD_BIT = 0x40000000u, HOT_SENSOR_BIT = 0x80000000u;
__host__ __device__ inline uint32_t hot_word(uint32_t slot, uint32_t cflags) {
    return slot | ((cflags & CF_SENSOR) ? HOT_SENSOR_BIT : 0u) | ((cflags & CF_COLD) ? HOT_COLD_BIT : 0u);
}
constexpr int32_t BODY_NO_COLLIDER = -1;  // body_col[] encoding; <= -2 : multi-collider body (handled by k_multi)

// Broadphase record of an active collider, as seen by the narrowphase. It is stored as two 16-byte halves:
//   hot  = (x, y, r, slot | needs_cold << 30 | is_sensor << 31) — cell-sorted array rebuilt every substep; enough for the
//                                                 self test and the distance prefilter
//   cold = (m, memberships, filter, parent)    — static per-collider array (ccold[slot]); fetched only for prefilter survivors
//                                                 whose record is flagged needs_cold. The default sphere (sole collider of its
//                                                 body, no mass override, not a sensor, groups ALL) has m = 2*(2r) = 4r exactly
//                                                 (collider.rs:40-42 + the doubled handle), so its cold half is synthesised.
struct Rec {
    float x, y, r, m;          // snapshot translation (physics.rs:360-366), radius, parent body's calculated_mass
    uint32_t memb, filt;       // InteractionGroups (groups.rs:7-12)
    uint32_t parent;           // parent body slot
    uint32_t slot_sensor;      // hot.w (see hot_word)
};

struct GridDesc {
    uint32_t W, H;       // toroidal table dims (cells)
    uint32_t ncells;     // W * H (per batched world)
    uint32_t n_worlds;   // batched independent worlds share one table: index = world * ncells + cell
    float cell;          // broadphase cell edge
    float inv_cell;      // 1 / cell (binning uses the monotone map floor(v * inv_cell))
    float rmax;          // max radius over active colliders (search reach = r + rmax)
    unsigned long long MW, MH;   // Lemire fastmod magics: 2^64 / W + 1, 2^64 / H + 1
};

struct Constraints {       // lib.rs:189-193: circle constraints, (x, y, radius, -) each, in global memory
    int n;
    const float4* c;
};

struct DeviceStats {       // accumulated per blobs_step* call, read back once
    unsigned long long collisions;
    unsigned long long coincident;
    unsigned long long rec_dropped;
    unsigned int nan_flag;
    unsigned int list_overflow;
    int bb_min_x, bb_min_y, bb_max_x, bb_max_y;   // bbox of collider snapshot cells (k_bbox)
    unsigned int max_ghosts, max_migrants;        // strip mode: largest message sections received in this call
    unsigned int over_count[2];                   // crowded-body list lengths, double-buffered by substep parity (k_crowded)
};

// Scene-query filter on the device (QueryFilter, query_filter.rs:74-108; flag values query_filter.rs:6-25)
struct QueryFilterDev {
    uint32_t flags;          // QueryFilterFlags bits
    uint32_t has_groups, memb, filt;
    uint32_t exclude_col;    // collider slot or NO_SLOT
    uint32_t exclude_body;   // body slot or NO_SLOT
    uint32_t wbase;          // first table entry of the batched world the query runs in
};

struct SubstepParams {
    float dt;
    float ratio_first, ratio_rest;  // dt/old_dt for the first non-static body, dt/dt for the rest (physics.rs:338-339, Q2)
    float gx, gy;
    uint32_t collisions_enabled;
    uint32_t n_bodies;              // body slots
    uint32_t n_colliders;           // collider slots
    uint32_t write_vel;             // materialise calculated_velocity this substep
    uint32_t crowded;               // bodies whose contact list overflows are deferred to k_crowded (else resolved inline)
    uint32_t over_parity;           // which DeviceStats::over_count entry this substep appends to
    uint32_t pool_min;              // (unused since the cooperative gather; kept for the parameter's ABI)
    uint32_t acc_zero;              // every body's acceleration is zero and no velocity_request is pending (any substep after the first of a call
                                    // in a world without springs: update_objects consumed both, physics.rs:334-336,350-355): neither is read or rewritten
    uint32_t nl_tail_decide;        // list pipeline: k_step is this substep's only publisher, so its last CTA decides for the next substep
    uint32_t nl_tail_publish;       // list pipeline on strips: ... and its last CTA publishes this rank's end-of-substep flag to every rank (no k_nls_publish launch)
    unsigned long long nl_cond_next; // ... and, inside a captured graph, tells the IF node that wraps the next substep's rebuild kernels (0 = none)
    uint32_t* over_list;            // deferred bodies: body slot, or OVER_MULTI_BIT | index into the multi-collider body list
};

struct BodyArrays {
    // device-authoritative state
    float2* pos;
    float2* pos_old;
    float2* acc;
    float2* vel;
    float2* vreq;
    uint8_t* has_vreq;
    float* rot;
    float* angvel;
    float* torque;
    // host-authoritative, packed so that a body needs two 8-byte loads
    const float* inertia;
    const uint2* binfo;      // (BF_* flags, body_col)
    const float2* bmg;       // (calculated_mass, gravity_mod)
    const uint32_t* bworld;  // batched-world id per body (only read when GridDesc::n_worlds > 1)
};

struct ColliderArrays {
    float2* cabs;            // snapshot translation (device-authoritative)
    uint2* ccell;            // (cell index, rank within cell) for the next table
    // host-authoritative
    const float2* coff;      // offset.translation (only read when CF_OFFSET is set)
    const uint4* cconst;     // (radius bits, CF_* flags, memberships, filter)
    const uint32_t* cparent; // body slot
    const uint4* ccold;      // (parent's calculated_mass bits, memberships, filter, parent slot) — Rec cold half
};

// ---- neighbour-list pipeline: contacts are found from per-collider lists of every collider within
// r_a + r_b + skin, rebuilt from the cell grid only when the snapshots have moved by more than ~skin/2 since the last build --
constexpr int NL_CAP = 32;                       // list rows; entry k of collider c sits at idx[k * stride + c] (transposed: coalesced)
constexpr int NL_SPEC = 4;                       // rows fetched together with the body state (before the count is known)
constexpr uint32_t NL_OVER = 0xffffffffu;        // NlView::hdr[c].z: more than NL_CAP neighbours -> the body goes to k_crowded
constexpr uint32_t NL_INACTIVE = 0xfffffffeu;    // NlView::hdr[c].z: collider slot is free / parentless / not owned by this rank
// NlView::hdr[c].w = CF_* flags | (strip-decomposed world) which neighbour rank keeps this collider as a ghost: every new snapshot
// record of such a collider is also stored straight into that neighbour's slot-indexed array (peer memory, NVLink)
constexpr uint32_t NLF_PUSH_L = 1u << 8, NLF_PUSH_R = 1u << 9;
constexpr int NL_MAX_RANKS = 16;
struct NlCtl {                 // device-resident control block
    unsigned int need;         // the lists are rebuilt before this substep's contact pass (decided by k_nl_decide)
    unsigned int force;        // host request: topology / table / staged collider writes changed
    unsigned int parity;       // which of the two tables holds the grid of the last rebuild
    unsigned int max_m;        // float bits: max over colliders of |snapshot - ref - c| (+ rounding slack), accumulated by the publishers
    float sum_x, sum_y;        // sampled sum of (snapshot - ref), and how many colliders were sampled
    unsigned int n_sum;
    float cx, cy;              // common-mode displacement the publishers of this substep subtract (any value is valid: it only decides WHEN lists are rebuilt)
    float mean_x, mean_y;      // sampled mean displacement at the previous decision (0 right after a rebuild)
    unsigned int done;         // CTA arrival counter of k_nl_build
    unsigned int step_done;    // CTA arrival counter of k_step (its last CTA takes the decision for the next substep)
    unsigned int decided;      // the decision for the coming substep has been taken already (by k_step's last CTA)
    unsigned int pub_seq;      // strip-decomposed world: sequence number of this rank's last k_nls_publish
    unsigned int published;    //                         ... and whether there has been one since the lists were (re)initialised
    unsigned long long rebuilds, substeps;   // statistics
};
struct NlView {
    const float4* snap_cur;    // slot-indexed snapshot records (x, y, r, hot word) of the previous substep: what the contact pass reads
    float4* snap_next;         // ... written by the publishers of this substep; nullptr = grid pipeline (no lists)
    uint4* hdr;                // per collider slot: (ref.x, ref.y, list length | NL_OVER | NL_INACTIVE, CF_* flags)
    uint32_t* idx;             // neighbour collider slots, ascending (= the reference's pair-loop order for a single-collider body)
    uint32_t stride;
    float lim;                 // rebuild when max_m exceeds this (0.45 * skin; +inf while collisions are disabled)
    float skin;
    NlCtl* ctl;
    uint32_t* tab[2];          // the two cell-start tables / scan-tile totals; [ctl->parity] is current
    uint32_t* tile[2];
    float4* hot;               // cell-sorted records of the last rebuild (their POSITIONS are stale: only the slot word is used)
    // strip-decomposed world (nullptr / unused otherwise)
    float4* peer_next[2];      // the left / right neighbour rank's snap_next (CUDA IPC mapping)
    const uint32_t* olist;     // compact list of the body slots this rank owns; k_step enumerates it
    const uint32_t* ocount;
};
// What a rank tells every other rank at the end of a substep (list pipeline on strips): its displacement-tracking accumulators.
// Every rank then takes the SAME rebuild decision from the same numbers (k_nls_decide).
struct NlFlag { unsigned int max_m; float sum_x, sum_y; unsigned int n_sum; unsigned int seq; unsigned int pad[3]; };
struct NlStripDev {
    int rank, nranks;
    NlFlag* mine;                  // [2][NL_MAX_RANKS]: slot [seq & 1][r] = what rank r published with sequence number seq
    NlFlag* peer[NL_MAX_RANKS];    // the same block on every rank (peer[rank] == mine)
    char* recv_block;              // rebuild-time ghost / migrant exchange: my receive block (k_strip_push layout: [side][parity] messages)
    char* peer_block[2];           // ... and the neighbours'
    size_t stride;
    unsigned int* xseq;            // exchange sequence number and CTA arrival counter of that exchange
    unsigned int* push_done;
};

struct Broadphase {
    const float4* hot;       // current cell-sorted hot halves (see Rec), read-only during the contact pass
    const uint32_t* tab;     // current cell starts, ncells + 1 entries
    uint32_t* tab_next;      // counts for the next table (zeroed)
    uint32_t* tile_next;     // per-scan-tile totals of tab_next (zeroed), so k_scan needs no inter-block dependency
    const float4* snap;      // list pipeline: records found through `hot` are re-read from this slot-indexed array (nullptr otherwise)
    NlView nl;
};

// ---- strip decomposition of one large world across ranks (BASELINE config #5) -------------------------------------
struct StripDesc {
    float x_lo, x_hi;          // this rank owns bodies whose collider snapshot x lies in [x_lo, x_hi)
    int has_left, has_right;
    float rmax;
    uint32_t gcap, mcap;       // capacities of the ghost / migration sections of a message
};
// full body state of a sphere that crossed a strip edge (the arrays are indexed by GLOBAL slot on every rank, so a
// migration is a plain write at [slot] on the receiving side — no device-side allocation)
struct MigRec {
    uint32_t slot, col;
    float2 pos, pos_old, acc, vel, vreq, cabs;
    float rot, angvel, torque;
    uint32_t has_vreq;
    uint32_t pad[2];
};
static_assert(sizeof(MigRec) == 80, "MigRec layout");
// message = header | float4 ghost[gcap] | MigRec mig[mcap]; fixed size so that no count has to reach the host
struct StripHeader { uint32_t n_ghost, n_mig, overflow, pad; };
__host__ __device__ inline size_t strip_msg_bytes(uint32_t gcap, uint32_t mcap) { return sizeof(StripHeader) + (size_t)gcap * 16 + (size_t)mcap * sizeof(MigRec); }
__host__ __device__ inline float4* strip_ghosts(void* msg) { return reinterpret_cast<float4*>(reinterpret_cast<char*>(msg) + sizeof(StripHeader)); }
__host__ __device__ inline MigRec* strip_migs(void* msg, uint32_t gcap) { return reinterpret_cast<MigRec*>(reinterpret_cast<char*>(msg) + sizeof(StripHeader) + (size_t)gcap * 16); }

// per-launch view of the strip state handed to k_main / k_scatter (all null / zero when strips are off)
struct StripView {
    const uint32_t* olist;     // compact list of owned body slots (NO_SLOT = released entry); nullptr = strips off
    const uint32_t* ocount;    // number of entries in olist (device-resident: arrivals are appended on the device)
    void* send_l;
    void* send_r;
    StripDesc S;
    const uint8_t* cowned;     // per collider slot: owned by this rank (k_tile enumerates records, ghosts included)
};

struct Recording {           // optional pair/event output
    uint32_t mode;           // 0 off, 1 pairs, 2 events
    uint32_t cap;
    unsigned long long* count;
    uint2* pairs;            // (slot_a, slot_b), a > b
    float4* vels;            // events: (vel_a.xy, vel_b.xy)
};

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr int SCAN_TILE_SHIFT = 13;
static_assert((1 << SCAN_TILE_SHIFT) == SCAN_TILE, "tile shift");

struct SpringParams { uint32_t a, b; float rest, k, c; };

struct JointParams { uint32_t a, b; float aax, aay, abx, aby, distance, target; };

// staged host writes to device-authoritative body state (insert_rbd / get_mut_rbd mutations)
struct BodyWrite {
    uint32_t slot, mask;
    float2 pos, pos_old, acc, vel, vreq;
    float rot, angvel, torque;
    uint32_t has_vreq;
};
enum : uint32_t { BW_POS = 1, BW_POS_OLD = 2, BW_ACC = 4, BW_VEL = 8, BW_VREQ = 16, BW_ROT = 32, BW_ANGVEL = 64, BW_TORQUE = 128,
                  BW_TRANSLATE = 256, BW_ADD_ACC = 512 };

struct ColWrite { uint32_t slot; float2 cabs; };

}  // namespace blobs
