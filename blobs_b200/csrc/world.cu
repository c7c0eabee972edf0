// Host orchestration of the GPU-resident world. See world.hpp / kernels.cuh.
#include <cstdlib>
#include "world.hpp"
#include "kernels.cuh"
#include "perf.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <numeric>

#include <dlfcn.h>
#include <nccl.h>

namespace blobs {

static void (*g_nccl_destroy)(void*) = nullptr;  // set once NCCL is loaded

#define CU(call)                                         \
    do {                                                 \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

// Kernel launch. The indirection exists for the host-compiled test build of this file (tests/emu, -DBLOBS_EMU), where
// `<<< >>>` is not C++; in the CUDA build it expands to the plain launch statement.
#ifdef BLOBS_EMU
#define BLOBS_LAUNCH(g, b, s, st, ...) ::emu::make_launch((g), (b), (s), __VA_ARGS__)
#else
#define BLOBS_LAUNCH(g, b, s, st, ...) __VA_ARGS__<<<(g), (b), (s), (st)>>>
#endif

// k_crowded runs a fixed grid whose warps stride over the list of deferred bodies (12 CTAs of 2 warps per SM)
#ifdef BLOBS_EMU
constexpr unsigned CROWD_GRID = 6;   // a fiber per thread: keep the empty CTAs few
#else
constexpr unsigned CROWD_GRID = 148 * 12;
#endif

static inline uint32_t h_slot(uint64_t h) { return (uint32_t)h; }
static inline unsigned cdiv(size_t a, unsigned b) { return (unsigned)((a + b - 1) / b); }

// device allocation that another process can map. Real CUDA: any cudaMalloc block has an IPC handle; the host-compiled test
// build needs memory shared between its rank processes.
static cudaError_t blobs_ipc_alloc(char** p, size_t bytes) {
#ifdef BLOBS_EMU
    return ::emu::ipc_alloc(reinterpret_cast<void**>(p), bytes);
#else
    return cudaMalloc(p, bytes);
#endif
}
static void blobs_ipc_free(char* p) {
#ifdef BLOBS_EMU
    ::emu::ipc_free(p);
#else
    cudaFree(p);
#endif
}

int World::cuda_fail(cudaError_t e, const char* what) {
    err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return BLOBS_ERR_CUDA;
}

World::World(const BlobsParams& p) : params(p) {
    gx = p.gravity.x;
    gy = p.gravity.y;
    use_spatial_hash = p.use_spatial_hash != 0;
    // test aids: initial values of the execution knobs (none of them changes results), so that a whole test-suite run can be
    // pushed through a non-default kernel path. Same meaning as the BLOBS_PARAM_* of the same name.
    if (const char* e = std::getenv("BLOBS_B200_POOL")) pool_mode = std::atoi(e);
    if (const char* e = std::getenv("BLOBS_B200_POOL_MIN")) pool_min = (uint32_t)std::atoi(e);
    if (const char* e = std::getenv("BLOBS_B200_CROWDED")) crowded_mode = std::atoi(e);
    if (const char* e = std::getenv("BLOBS_B200_TUNE")) tune = std::atoi(e);
    if (const char* e = std::getenv("BLOBS_B200_LIST")) list_mode = std::atoi(e);
    if (const char* e = std::getenv("BLOBS_B200_COND")) cond_nodes = std::atoi(e) != 0;
    if (const char* e = std::getenv("BLOBS_B200_SKIN")) skin_frac = (float)std::atof(e);
    if (const char* e = std::getenv("BLOBS_B200_STRIP_P2P")) p2p_request = std::atoi(e) != 0;
    if (const char* e = std::getenv("BLOBS_B200_STRIP_GRAPH")) strip_graph = std::atoi(e) != 0;
    if (const char* e = std::getenv("BLOBS_B200_NLS_TAIL")) nls_tail_publish = std::atoi(e) != 0;
    if (const char* e = std::getenv("BLOBS_B200_JADV")) joint_advance = std::atoi(e) != 0;
#ifdef BLOBS_EMU
    graphs_on = false;   // host-compiled test build (tests/emu): no CUDA graphs there
#endif
}

int World::init() {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(BLOBS_ERR_CUDA, std::string("no usable CUDA device (libblobs_b200 has no CPU fallback): ") +
                                        (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (params.device >= 0) {
        CU(cudaSetDevice(params.device));
        device = params.device;
    } else {
        CU(cudaGetDevice(&device));
    }
    CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CU(cudaMalloc(&d_stats, sizeof(DeviceStats)));
    CU(cudaMallocHost(&h_stats, sizeof(DeviceStats)));
    CU(cudaMalloc(&d_nlctl, sizeof(NlCtl)));
    CU(cudaMemsetAsync(d_nlctl, 0, sizeof(NlCtl), stream));
    CU(cudaMallocHost(&h_nlctl, sizeof(NlCtl)));
    std::memset(h_nlctl, 0, sizeof(NlCtl));
    CU(cudaMalloc(&d_rec_count, sizeof(unsigned long long)));
    CU(cudaMemsetAsync(d_rec_count, 0, sizeof(unsigned long long), stream));
    CU(cudaEventCreate(&ev_step0));
    CU(cudaEventCreate(&ev_step1));
    CU(cudaFuncSetAttribute(k_joints_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, JOINT_THREADS * JOINT_SMEM_MAX * (int)sizeof(float4)));
    if (params.body_capacity_hint) {
        const size_t n = params.body_capacity_hint;
        CU(pos.ensure(n, stream)); CU(pos_old.ensure(n, stream)); CU(acc.ensure(n, stream)); CU(vel.ensure(n, stream));
        CU(vreq.ensure(n, stream)); CU(has_vreq.ensure(n, stream)); CU(rot.ensure(n, stream)); CU(angvel.ensure(n, stream));
        CU(torque.ensure(n, stream));
    }
    if (params.collider_capacity_hint) {
        const size_t n = params.collider_capacity_hint;
        CU(cabs.ensure(n, stream)); CU(ccell.ensure(n, stream));
    }
    return BLOBS_OK;
}

World::~World() {
    // strip mode: the stream may be parked inside a collective whose peer is gone — never block process exit on it
    if (stream && !strip_on) cudaStreamSynchronize(stream);
    if (s_body) cudaStreamDestroy(s_body);
    if (io_ready) { cudaStreamSynchronize(s_h2d); cudaStreamDestroy(s_h2d); cudaStreamSynchronize(s_d2h); cudaStreamDestroy(s_d2h); }
    for (cudaEvent_t e : {ev_up_done[0], ev_up_done[1], ev_up_free[0], ev_up_free[1], ev_snap_ready, ev_snap_free}) if (e) cudaEventDestroy(e);
    d_forces_up[0].release(); d_forces_up[1].release(); d_pos_snap.release();
    d_fslots_up[0].release(); d_fslots_up[1].release(); d_oslots_snap.release(); d_oxy_snap.release();
    if (d_ocount_snap) cudaFree(d_ocount_snap);
    for (auto& e : ev_pool) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    destroy_graph(gslot[0]);
    destroy_graph(gslot[1]);
    if (ev_step0) cudaEventDestroy(ev_step0);
    if (ev_step1) cudaEventDestroy(ev_step1);
    inertia.d.release(); binfo.d.release(); bmg.d.release(); bworld.d.release();
    coff.d.release(); cconst.d.release(); cparent.d.release(); ccold.d.release();
    pos.release(); pos_old.release(); acc.release(); vel.release(); vreq.release(); cabs.release();
    has_vreq.release(); rot.release(); angvel.release(); torque.release(); ccell.release();
    d_pending.release(); d_pending_col.release();
    mb_body.release(); mb_off.release(); mb_cols.release(); sb_body.release(); sb_off.release(); sb_edge.release();
    isl_off.release(); isl_joint.release(); d_joints_inter.release(); isl_boff.release(); isl_body.release(); d_springs.release(); d_joints.release();
    hot_a.release(); hot_b.release(); tab_a.release(); tab_b.release(); tile_a.release(); tile_b.release(); over_list.release();
    d_qcentre.release(); d_qradius.release(); d_qcount.release(); d_qoff.release(); d_qhits.release();
    rec_pairs.release(); rec_vels.release(); d_sub_end.release(); d_forces.release(); d_constraints.release(); d_cellx.release(); d_celly.release();
    for (int i = 0; i < 4; ++i) if (msg[i]) cudaFree(msg[i]);
    for (int i = 0; i < 2; ++i) if (p2p_peer[i]) cudaIpcCloseMemHandle(p2p_peer[i]);
    if (p2p_block) blobs_ipc_free(p2p_block);
    if (d_push_done) cudaFree(d_push_done);
    d_owned.release(); d_cowned.release(); gcell.release(); io_slots.release(); io_xy.release(); olist.release(); opos.release();
    if (d_ocount) cudaFree(d_ocount);
    if (d_io_count) cudaFree(d_io_count);
    if (nccl_comm && g_nccl_destroy) g_nccl_destroy(nccl_comm);
    if (nls_snap_block) { snap_a.d = snap_b.d = nullptr; snap_a.cap = snap_b.cap = 0; blobs_ipc_free(nls_snap_block); }
    for (int i = 0; i < 2; ++i) if (nls_peer_snap[i]) cudaIpcCloseMemHandle(nls_peer_snap[i]);
    for (int r = 0; r < NL_MAX_RANKS; ++r) if (nls_peer_flags[r] && nls_peer_flags[r] != nls_flags) cudaIpcCloseMemHandle(nls_peer_flags[r]);
    if (nls_flags) blobs_ipc_free(reinterpret_cast<char*>(nls_flags));
    snap_a.release(); snap_b.release(); nl_hdr.release(); nl_idx.release();
    if (d_nlctl) cudaFree(d_nlctl);
    if (h_nlctl) cudaFreeHost(h_nlctl);
    if (d_stats) cudaFree(d_stats);
    if (h_stats) cudaFreeHost(h_stats);
    if (d_rec_count) cudaFree(d_rec_count);
    if (stream) cudaStreamDestroy(stream);
}

// ---------------------------------------------------------------------------------------------- params
int World::set_param(int id, double v) {
    switch (id) {
        case BLOBS_PARAM_GRAVITY_X: gx = (float)v; break;
        case BLOBS_PARAM_GRAVITY_Y: gy = (float)v; break;
        case BLOBS_PARAM_SUBSTEPS: substeps = (uint32_t)v; break;
        case BLOBS_PARAM_JOINT_ITERATIONS: joint_iterations = (uint32_t)v; break;
        case BLOBS_PARAM_USE_SPATIAL_HASH: use_spatial_hash = v != 0; break;
        case BLOBS_PARAM_COLLISIONS_ENABLED: collisions_enabled = v != 0; break;
        case BLOBS_PARAM_ACCUMULATOR: accumulator = v; break;
        case BLOBS_PARAM_TIME: time = v; break;
        case BLOBS_PARAM_OLD_DT: old_dt = (float)v; break;
        case BLOBS_PARAM_CELL_SIZE: cell_size = (float)v; break;
        case BLOBS_PARAM_BROADPHASE_CELL: bp_cell_override = (float)v; bp_dirty = true; break;
        case BLOBS_PARAM_CONTACT_MODE: contact_mode = (int)v; break;
        case BLOBS_PARAM_FUSED: allow_fused = v != 0; break;
        case BLOBS_PARAM_TUNE: tune = (int)v; break;
        case BLOBS_PARAM_CROWDED: crowded_mode = (int)v; break;
        case BLOBS_PARAM_POOL: pool_mode = (int)v; break;
        case BLOBS_PARAM_POOL_MIN: pool_min = (uint32_t)v; break;
        case BLOBS_PARAM_STRIP_P2P:
            if (strip_on) return fail(BLOBS_ERR_INVALID, "BLOBS_PARAM_STRIP_P2P must be set before blobs_strip_configure");
            p2p_request = v != 0;
            break;
        case BLOBS_PARAM_GRAPH: graphs_on = v != 0; break;
        case BLOBS_PARAM_LIST: list_mode = (int)v; nl_grid_hold = 0; nl_next_hold = 32; bp_dirty = true; break;
        case BLOBS_PARAM_SKIN:
            if (!(v > 0.0) || !(v <= 16.0)) return fail(BLOBS_ERR_INVALID, "BLOBS_PARAM_SKIN must be in (0, 16] (fraction of the largest collider radius)");
            skin_frac = (float)v; bp_dirty = true;
            break;
        case BLOBS_PARAM_STRIP_MAX_GHOSTS: last_max_ghosts = (uint32_t)v; break;      // reset
        case BLOBS_PARAM_STRIP_MAX_MIGRANTS: last_max_migrants = (uint32_t)v; break;  // reset
        case BLOBS_PARAM_BATCH_WORLD:
            if (v < 0 || v >= 1048576.0) return fail(BLOBS_ERR_INVALID, "batch world id out of range");
            cur_world = (uint32_t)v;
            break;
        default: return fail(BLOBS_ERR_INVALID, "unknown param id");
    }
    return BLOBS_OK;
}

int World::get_param(int id, double* out) const {
    switch (id) {
        case BLOBS_PARAM_GRAVITY_X: *out = gx; break;
        case BLOBS_PARAM_GRAVITY_Y: *out = gy; break;
        case BLOBS_PARAM_SUBSTEPS: *out = substeps; break;
        case BLOBS_PARAM_JOINT_ITERATIONS: *out = joint_iterations; break;
        case BLOBS_PARAM_USE_SPATIAL_HASH: *out = use_spatial_hash; break;
        case BLOBS_PARAM_COLLISIONS_ENABLED: *out = collisions_enabled; break;
        case BLOBS_PARAM_ACCUMULATOR: *out = accumulator; break;
        case BLOBS_PARAM_TIME: *out = time; break;
        case BLOBS_PARAM_OLD_DT: *out = old_dt; break;
        case BLOBS_PARAM_CELL_SIZE: *out = cell_size; break;
        case BLOBS_PARAM_BROADPHASE_CELL: *out = bp_cell_override; break;
        case BLOBS_PARAM_CONTACT_MODE: *out = contact_mode; break;
        case BLOBS_PARAM_FUSED: *out = allow_fused; break;
        case BLOBS_PARAM_TUNE: *out = tune; break;
        case BLOBS_PARAM_CROWDED: *out = crowded_mode; break;
        case BLOBS_PARAM_POOL: *out = pool_mode; break;
        case BLOBS_PARAM_POOL_MIN: *out = pool_min; break;
        case BLOBS_PARAM_STRIP_P2P: *out = strip_on ? (p2p_on ? 1.0 : 0.0) : (p2p_request ? 1.0 : 0.0); break;
        case BLOBS_PARAM_BATCH_WORLD: *out = cur_world; break;
        case BLOBS_PARAM_GRAPH: *out = graphs_on; break;
        case BLOBS_PARAM_LIST: *out = list_mode; break;
        case BLOBS_PARAM_LIST_ACTIVE: *out = nl_on ? 1.0 : 0.0; break;
        case BLOBS_PARAM_SKIN: *out = skin_frac; break;
        case BLOBS_PARAM_LIST_REBUILDS: *out = (double)h_nlctl->rebuilds; break;
        case BLOBS_PARAM_LIST_SUBSTEPS: *out = (double)h_nlctl->substeps; break;
        case BLOBS_PARAM_GRAPH_REPLAYS: *out = (double)graph_replays; break;
        case BLOBS_PARAM_STRIP_MAX_GHOSTS: *out = last_max_ghosts; break;
        case BLOBS_PARAM_STRIP_MAX_MIGRANTS: *out = last_max_migrants; break;
        default: return BLOBS_ERR_INVALID;
    }
    return BLOBS_OK;
}

// Physics::reset (physics.rs:71-76): clears the four arenas; constraints, time, old_dt stay.
int World::reset() {
    bodies.clear(); cols.clear(); springs.clear(); joints.clear();
    for (auto& b : hb) b = HBody{};
    pending.clear();
    std::fill(pending_idx.begin(), pending_idx.end(), -1);
    pending_col.clear();
    topo_dirty = bp_dirty = true;
    shadow_valid = false;
    n_worlds = 1;        // the batched-world partition goes with the bodies
    cur_world = 0;
    return BLOBS_OK;
}

// ---------------------------------------------------------------------------------------------- staging
BodyWrite& World::stage(uint32_t slot) {
    if (pending_idx.size() < bodies.slots()) pending_idx.resize(bodies.slots(), -1);
    int32_t i = pending_idx[slot];
    if (i < 0) {
        i = (int32_t)pending.size();
        pending_idx[slot] = i;
        BodyWrite w{};
        w.slot = slot;
        pending.push_back(w);
    }
    return pending[i];
}

int World::ensure_capacity() {
    const size_t nb = bodies.slots(), nc = cols.slots();
    CU(pos.ensure(nb, stream)); CU(pos_old.ensure(nb, stream)); CU(acc.ensure(nb, stream)); CU(vel.ensure(nb, stream));
    CU(vreq.ensure(nb, stream)); CU(has_vreq.ensure(nb, stream)); CU(rot.ensure(nb, stream)); CU(angvel.ensure(nb, stream));
    CU(torque.ensure(nb, stream)); CU(cabs.ensure(nc, stream)); CU(ccell.ensure(nc, stream));
    CU(over_list.ensure(nb, stream));
    return BLOBS_OK;
}

int World::flush_writes() {
    if (!pending.empty() || !pending_col.empty()) {
        int rc = ensure_capacity();
        if (rc) return rc;
    }
    if (!pending.empty()) {
        CU(d_pending.ensure(pending.size(), stream));
        CU(cudaMemcpyAsync(d_pending.d, pending.data(), pending.size() * sizeof(BodyWrite), cudaMemcpyHostToDevice, stream));
        BLOBS_LAUNCH(cdiv(pending.size(), 256), 256, 0, stream, k_apply_body_writes)(body_arrays(), d_pending.d, (uint32_t)pending.size());
        launches++;
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(stream));  // pending is pageable host memory
        for (auto& w : pending) pending_idx[w.slot] = -1;
        pending.clear();
    }
    if (!pending_col.empty()) {
        nl_force_pending = true;   // caller-supplied snapshots: the lists (and the slot-indexed records) must be rebuilt from them
        CU(d_pending_col.ensure(pending_col.size(), stream));
        CU(cudaMemcpyAsync(d_pending_col.d, pending_col.data(), pending_col.size() * sizeof(ColWrite), cudaMemcpyHostToDevice, stream));
        BLOBS_LAUNCH(cdiv(pending_col.size(), 256), 256, 0, stream, k_apply_col_writes)(cabs.d, d_pending_col.d, (uint32_t)pending_col.size());
        launches++;
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(stream));
        pending_col.clear();
    }
    return BLOBS_OK;
}

BodyArrays World::body_arrays() {
    BodyArrays B;
    B.pos = pos.d; B.pos_old = pos_old.d; B.acc = acc.d; B.vel = vel.d; B.vreq = vreq.d; B.has_vreq = has_vreq.d;
    B.rot = rot.d; B.angvel = angvel.d; B.torque = torque.d; B.inertia = inertia.d.d; B.binfo = binfo.d.d; B.bmg = bmg.d.d; B.bworld = bworld.d.d;
    return B;
}
ColliderArrays World::col_arrays() {
    ColliderArrays C;
    C.cabs = cabs.d; C.ccell = ccell.d; C.coff = coff.d.d; C.cconst = cconst.d.d; C.cparent = cparent.d.d; C.ccold = ccold.d.d;
    return C;
}
Constraints World::constraints_pod() const {
    Constraints K;
    K.n = (int)con_pos.size();
    K.c = d_constraints.d;
    return K;
}

// ---------------------------------------------------------------------------------------------- bodies
int World::body_insert(const BlobsBodyDesc& d, uint64_t* out) {
    const uint64_t h = bodies.insert();
    const uint32_t s = h_slot(h);
    if (hb.size() < bodies.slots()) hb.resize(bodies.slots());
    HBody& b = hb[s];
    b = HBody{};
    b.ud_lo = d.user_data_lo; b.ud_hi = d.user_data_hi;
    b.scale = d.scale;
    b.type = d.body_type;
    b.rot_active = d.rotation != 0.0f;
    b.world = cur_world;
    if (cur_world + 1 > n_worlds) { n_worlds = cur_world + 1; bp_dirty = true; }
    const size_t n = bodies.slots();
    inertia.resize(n, 1.0f); bmg.resize(n, make_float2(1.0f, 1.0f)); binfo.resize(n, make_uint2(0u, (uint32_t)BODY_NO_COLLIDER));
    bworld.resize(n, 0u);
    bworld.set(s, cur_world);
    bmg.set(s, make_float2(1.0f, d.gravity_mod));  // calculated_mass = 1: RigidBodyBuilder::build rigid_body.rs:385-388
    inertia.set(s, 1.0f);
    BodyWrite& w = stage(s);
    w.mask = BW_POS | BW_POS_OLD | BW_ACC | BW_VEL | BW_VREQ | BW_ROT | BW_ANGVEL | BW_TORQUE;
    w.pos = make_float2(d.position.x, d.position.y);
    w.pos_old = make_float2(d.position_old.x, d.position_old.y);
    w.acc = make_float2(d.acceleration.x, d.acceleration.y);
    w.vel = make_float2(d.calculated_velocity.x, d.calculated_velocity.y);
    w.vreq = make_float2(d.velocity_request.x, d.velocity_request.y);
    w.has_vreq = d.has_velocity_request ? 1u : 0u;
    w.rot = d.rotation; w.angvel = 0.f; w.torque = 0.f;
    if (shadow_valid) {
        if (sh_pos.size() < n) { sh_pos.resize(n); sh_rot.resize(n); }
        sh_pos[s] = w.pos;
        sh_rot[s] = w.rot;
    }
    topo_dirty = true;
    if (out) *out = h;
    return BLOBS_OK;
}

// rigid_body.rs:96-128
void World::update_mass_and_inertia(uint32_t bs) {
    HBody& b = hb[bs];
    float m = 0.0f, in = 0.0f;
    float wx = 0.0f, wy = 0.0f;
    for (uint64_t ch : b.colliders) {
        if (!cols.valid(ch)) continue;
        const BlobsColliderDesc& c = hc[h_slot(ch)].desc;
        if (c.is_sensor) continue;
        const float cm = c.has_mass_override ? c.mass_override : c.radius * 2.0f;   // collider.rs:40-42
        const float ox = c.offset.translation.x, oy = c.offset.translation.y;
        const float dlen = std::sqrt(ox * ox + oy * oy);
        const float ci = 0.5f * cm * (c.radius * c.radius);                         // collider.rs:44-50
        m += cm;
        in += ci + cm * (dlen * dlen);
        wx += ox * cm;
        wy += oy * cm;
    }
    if (m == 0.0f) m = 1.0f;
    if (in == 0.0f) in = 1.0f;
    b.com = BlobsVec2{wx / m, wy / m};
    set_mass(bs, m);
    inertia.set(bs, in);
    topo_dirty = true;  // ccold[] carries the parent's mass
}

int World::body_remove(uint64_t h) {
    if (!bodies.valid(h)) {   // rigid_body.rs:266-275: not a panic, an Error entry in the event history
        PhysicsEventRec ev;
        ev.message = "removing a non-existent rigid body";
        ev.severity = 4;
        ev.rbd_handle = h;
        EventHistory::global().push(std::move(ev));
        return fail(BLOBS_ERR_STALE_HANDLE, "removing a non-existent rigid body");
    }
    const uint32_t s = h_slot(h);
    for (uint64_t ch : hb[s].colliders)
        if (cols.valid(ch)) cols.remove_slot(h_slot(ch));  // remove_ignoring_parent (physics.rs:165-167)
    hb[s] = HBody{};
    bodies.remove_slot(s);
    topo_dirty = bp_dirty = true;
    return BLOBS_OK;
}

// position of one body as the host would see it now (staged writes applied); no topology work
int World::peek_position(uint32_t slot, float2* out) {
    if (shadow_valid && slot < sh_pos.size()) { *out = sh_pos[slot]; return BLOBS_OK; }
    int rc = flush_writes();
    if (rc) return rc;
    if (slot >= pos.cap) return BLOBS_ERR_INVALID;
    CU(cudaMemcpyAsync(out, pos.d + slot, sizeof(float2), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    return BLOBS_OK;
}

int World::ensure_shadow() {
    if (shadow_valid) return BLOBS_OK;  // staged writes keep a valid shadow up to date (body_insert / body_set)
    int rc = flush_writes();            // only the device-authoritative state matters here: no topology rebuild
    if (rc) return rc;
    const size_t n = bodies.slots();
    sh_pos.resize(n);
    sh_rot.resize(n);
    if (n) {
        rc = ensure_capacity();
        if (rc) return rc;
        CU(cudaMemcpyAsync(sh_pos.data(), pos.d, n * sizeof(float2), cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(sh_rot.data(), rot.d, n * sizeof(float), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
    }
    shadow_valid = true;
    return BLOBS_OK;
}

int World::body_get(uint64_t h, BlobsBodyState* out) {
    if (!bodies.valid(h)) return fail(BLOBS_ERR_STALE_HANDLE, "get_rbd: None");
    int rc = flush();
    if (rc) return rc;
    const uint32_t s = h_slot(h);
    float2 p, po, a, v, vr;
    float r, w, t;
    uint8_t hv;
    CU(cudaMemcpyAsync(&p, pos.d + s, sizeof(float2), cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpyAsync(&po, pos_old.d + s, sizeof(float2), cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpyAsync(&a, acc.d + s, sizeof(float2), cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpyAsync(&v, vel.d + s, sizeof(float2), cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpyAsync(&vr, vreq.d + s, sizeof(float2), cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpyAsync(&r, rot.d + s, sizeof(float), cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpyAsync(&w, angvel.d + s, sizeof(float), cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpyAsync(&t, torque.d + s, sizeof(float), cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpyAsync(&hv, has_vreq.d + s, 1, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    const HBody& b = hb[s];
    std::memset(out, 0, sizeof(*out));
    out->position = {p.x, p.y}; out->position_old = {po.x, po.y}; out->acceleration = {a.x, a.y};
    out->calculated_velocity = {v.x, v.y}; out->velocity_request = {vr.x, vr.y}; out->has_velocity_request = hv;
    out->rotation = r; out->angular_velocity = w; out->torque = t;
    out->center_of_mass = b.com; out->scale = b.scale;
    out->calculated_mass = bmg.h[s].x; out->inertia = inertia.h[s]; out->gravity_mod = bmg.h[s].y;
    out->body_type = b.type; out->user_data_lo = b.ud_lo; out->user_data_hi = b.ud_hi;
    return BLOBS_OK;
}

int World::body_set(uint64_t h, const BlobsBodyState& s, uint32_t mask) {
    if (!bodies.valid(h)) return fail(BLOBS_ERR_STALE_HANDLE, "get_mut_rbd: None");
    const uint32_t slot = h_slot(h);
    HBody& b = hb[slot];
    const uint32_t dev_mask = BLOBS_BODY_POSITION | BLOBS_BODY_POSITION_OLD | BLOBS_BODY_ACCELERATION | BLOBS_BODY_VELOCITY_REQUEST |
                              BLOBS_BODY_CALC_VELOCITY | BLOBS_BODY_ROTATION | BLOBS_BODY_ANGULAR_VELOCITY | BLOBS_BODY_TORQUE;
    if (mask & dev_mask) {
        {
            BodyWrite& w0 = stage(slot);
            if (w0.mask & (BW_TRANSLATE | BW_ADD_ACC)) {  // keep read-modify-write ops ordered
                int rc = flush_writes();
                if (rc) return rc;
            }
        }
        BodyWrite& w = stage(slot);
        if (mask & BLOBS_BODY_POSITION) { w.mask |= BW_POS; w.pos = make_float2(s.position.x, s.position.y); if (shadow_valid) sh_pos[slot] = w.pos; }
        if (mask & BLOBS_BODY_POSITION_OLD) { w.mask |= BW_POS_OLD; w.pos_old = make_float2(s.position_old.x, s.position_old.y); }
        if (mask & BLOBS_BODY_ACCELERATION) { w.mask |= BW_ACC; w.acc = make_float2(s.acceleration.x, s.acceleration.y); }
        if (mask & BLOBS_BODY_CALC_VELOCITY) { w.mask |= BW_VEL; w.vel = make_float2(s.calculated_velocity.x, s.calculated_velocity.y); }
        if (mask & BLOBS_BODY_VELOCITY_REQUEST) {
            w.mask |= BW_VREQ;
            w.vreq = make_float2(s.velocity_request.x, s.velocity_request.y);
            w.has_vreq = s.has_velocity_request ? 1u : 0u;
        }
        if (mask & BLOBS_BODY_ROTATION) { w.mask |= BW_ROT; w.rot = s.rotation; if (shadow_valid) sh_rot[slot] = w.rot; if (s.rotation != 0.f) b.rot_active = true; }
        if (mask & BLOBS_BODY_ANGULAR_VELOCITY) { w.mask |= BW_ANGVEL; w.angvel = s.angular_velocity; if (s.angular_velocity != 0.f) b.rot_active = true; }
        if (mask & BLOBS_BODY_TORQUE) { w.mask |= BW_TORQUE; w.torque = s.torque; if (s.torque != 0.f) b.rot_active = true; }
        if (b.rot_active && !(binfo.h[slot].x & BF_ROT)) topo_dirty = true;
    }
    if (mask & BLOBS_BODY_MASS) { set_mass(slot, s.calculated_mass); topo_dirty = true; }
    if (mask & BLOBS_BODY_INERTIA) inertia.set(slot, s.inertia);
    if (mask & BLOBS_BODY_GRAVITY_MOD) set_gmod(slot, s.gravity_mod);
    if (mask & BLOBS_BODY_TYPE) { if (b.type != s.body_type) { b.type = s.body_type; topo_dirty = true; } }
    if (mask & BLOBS_BODY_USER_DATA) { b.ud_lo = s.user_data_lo; b.ud_hi = s.user_data_hi; }
    if (mask & BLOBS_BODY_SCALE) b.scale = s.scale;
    if (mask & BLOBS_BODY_CENTER_OF_MASS) b.com = s.center_of_mass;
    return BLOBS_OK;
}

// update_rigid_body_position (physics.rs:174-182): position += offset
int World::body_translate(uint64_t h, BlobsVec2 off) {
    if (!bodies.valid(h)) return BLOBS_OK;  // `if let Some(..)`: silently ignored
    const uint32_t slot = h_slot(h);
    if (stage(slot).mask & (BW_TRANSLATE | BW_POS)) {
        int rc = flush_writes();
        if (rc) return rc;
    }
    BodyWrite& w = stage(slot);
    w.mask |= BW_TRANSLATE;
    w.pos = make_float2(off.x, off.y);
    shadow_valid = false;
    return BLOBS_OK;
}

// RigidBody::apply_force (rigid_body.rs:155-160)
int World::body_apply_force(uint64_t h, BlobsVec2 f) {
    if (!bodies.valid(h)) return fail(BLOBS_ERR_STALE_HANDLE, "apply_force: stale handle");
    const uint32_t slot = h_slot(h);
    if (hb[slot].type == BLOBS_BODY_STATIC) return BLOBS_OK;
    if (stage(slot).mask & (BW_ADD_ACC | BW_ACC)) {
        int rc = flush_writes();
        if (rc) return rc;
    }
    BodyWrite& w = stage(slot);
    w.mask |= BW_ADD_ACC;
    const float m = bmg.h[slot].x;
    w.acc = make_float2(f.x / m, f.y / m);
    return BLOBS_OK;
}

int World::body_colliders(uint64_t h, uint64_t* out, size_t cap, size_t* n) const {
    if (!bodies.valid(h)) return BLOBS_ERR_STALE_HANDLE;
    const auto& v = hb[h_slot(h)].colliders;
    for (size_t i = 0; i < v.size() && i < cap; ++i) out[i] = v[i];
    *n = v.size();
    return BLOBS_OK;
}

// ---------------------------------------------------------------------------------------------- colliders
int World::collider_insert(const BlobsColliderDesc& d, uint64_t parent, uint64_t* out) {
    if (!bodies.valid(parent)) return fail(BLOBS_ERR_STALE_HANDLE, "parent rigid body must exist when inserting collider");  // physics.rs:142
    const uint64_t h = cols.insert();
    const uint32_t s = h_slot(h);
    if (hc.size() < cols.slots()) hc.resize(cols.slots());
    hc[s].desc = d;
    hc[s].parent = parent;
    hc[s].born_epoch = snap_epoch;
    const size_t n = cols.slots();
    coff.resize(n, make_float2(0.f, 0.f)); cconst.resize(n, make_uint4(0u, 0u, 0u, 0u)); cparent.resize(n, NO_SLOT);
    ccold.resize(n, make_uint4(0u, 0u, 0u, NO_SLOT));
    coff.set(s, make_float2(d.offset.translation.x, d.offset.translation.y));
    {
        uint32_t rbits;
        std::memcpy(&rbits, &d.radius, 4);
        cconst.set(s, make_uint4(rbits, 0u, d.memberships, d.filter));  // flags are resolved in rebuild_topology
    }
    pending_col.push_back(ColWrite{s, make_float2(d.absolute_transform.translation.x, d.absolute_transform.translation.y)});
    HBody& b = hb[h_slot(parent)];
    b.colliders.push_back(h);  // collider.rs:179-181
    b.colliders.push_back(h);  // physics.rs:144
    b.cols.push_back(s);
    update_mass_and_inertia(h_slot(parent));  // physics.rs:146
    topo_dirty = bp_dirty = true;
    if (out) *out = h;
    return BLOBS_OK;
}

// remove_col (physics.rs:159-161 -> collider.rs:134-164)
int World::collider_remove(uint64_t h) {
    if (!cols.valid(h)) return fail(BLOBS_ERR_STALE_HANDLE, "remove_col: collider not found");
    const uint32_t s = h_slot(h);
    const uint64_t parent = hc[s].parent;
    if (parent != 0 && bodies.valid(parent)) {
        const uint32_t bs = h_slot(parent);
        HBody& b = hb[bs];
        b.colliders.erase(std::remove(b.colliders.begin(), b.colliders.end(), h), b.colliders.end());
        b.cols.erase(std::remove(b.cols.begin(), b.cols.end(), s), b.cols.end());
        update_mass_and_inertia(bs);
        if (b.colliders.empty()) {  // "rbd removed because colliders.len() == 0" (collider.rs:143-158)
            PhysicsEventRec ev;
            ev.message = "rbd removed because colliders.len() == 0";
            ev.severity = 2;
            ev.col_handle = h;
            ev.rbd_handle = parent;
            float2 p;
            if (peek_position(bs, &p) == BLOBS_OK) { ev.has_position = true; ev.px = p.x; ev.py = p.y; }
            EventHistory::global().push(std::move(ev));
            hb[bs] = HBody{};
            bodies.remove_slot(bs);
        }
    }
    cols.remove_slot(s);
    topo_dirty = bp_dirty = true;
    return BLOBS_OK;
}

int World::collider_get(uint64_t h, BlobsColliderState* out) {
    if (!cols.valid(h)) return fail(BLOBS_ERR_STALE_HANDLE, "get_col: None");
    int rc = flush();
    if (rc) return rc;
    const uint32_t s = h_slot(h);
    float2 a;
    std::vector<float> r(bodies.slots());
    CU(cudaMemcpyAsync(&a, cabs.d + s, sizeof(float2), cudaMemcpyDeviceToHost, stream));
    if (!r.empty()) CU(cudaMemcpyAsync(r.data(), rot.d, r.size() * sizeof(float), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    out->desc = hc[s].desc;
    out->desc.absolute_transform = live_snapshot(s, a, r);
    out->parent = hc[s].parent;
    return BLOBS_OK;
}

// ---------------------------------------------------------------------------------------------- springs / joints
int World::spring_insert(uint64_t a, uint64_t b, float rest, float k, float c, uint64_t* out) {
    const uint64_t h = springs.insert();
    if (hs.size() < springs.slots()) hs.resize(springs.slots());
    hs[h_slot(h)] = HSpring{a, b, rest, k, c};
    topo_dirty = true;
    if (out) *out = h;
    return BLOBS_OK;
}
int World::spring_remove(uint64_t h) {
    if (!springs.valid(h)) return BLOBS_ERR_STALE_HANDLE;
    springs.remove_slot(h_slot(h));
    topo_dirty = true;
    return BLOBS_OK;
}

int World::joint_insert(uint64_t a, uint64_t b, BlobsVec2 aa, BlobsVec2 ab, float dist, uint64_t* out) {
    if (h_slot(a) == h_slot(b) && bodies.valid(a) && bodies.valid(b)) return fail(BLOBS_ERR_SAME_BODY, "get2_mut called with identical indices");
    if (!bodies.valid(a) || !bodies.valid(b)) return fail(BLOBS_ERR_STALE_HANDLE, "create_fixed_joint: unwrap on None");
    int rc = ensure_shadow();
    if (rc) return rc;
    const uint32_t sa = h_slot(a), sb = h_slot(b);
    if (std::isnan(dist)) {  // create_fixed_joint (physics.rs:198)
        const float2 pa = sh_pos[sa], pb = sh_pos[sb];
        const float x = ((pa.x + aa.x) - pb.x) - ab.x, y = ((pa.y + aa.y) - pb.y) - ab.y;
        dist = std::sqrt(x * x + y * y);
    }
    const uint64_t h = joints.insert();
    if (hj.size() < joints.slots()) hj.resize(joints.slots());
    hj[h_slot(h)] = HJoint{a, b, aa, ab, dist, sh_rot[sb] - sh_rot[sa]};  // physics.rs:230
    hb[sa].joints.push_back(h);
    hb[sb].joints.push_back(h);
    topo_dirty = true;
    if (out) *out = h;
    return BLOBS_OK;
}
int World::joint_remove(uint64_t h) {
    if (!joints.valid(h)) return BLOBS_ERR_STALE_HANDLE;
    joints.remove_slot(h_slot(h));
    topo_dirty = true;
    return BLOBS_OK;
}

int World::constraint_push(BlobsVec2 p, float r) {
    con_pos.push_back(p);
    con_r.push_back(r);
    con_dirty = true;
    return BLOBS_OK;
}
int World::constraint_clear() {
    con_pos.clear();
    con_r.clear();
    con_dirty = true;
    return BLOBS_OK;
}

// ---------------------------------------------------------------------------------------------- topology
template <class T>
static cudaError_t upload(DevBuf<T>& d, const std::vector<T>& h, cudaStream_t st) {
    cudaError_t e = d.ensure(std::max<size_t>(h.size(), 1), st);
    if (e != cudaSuccess) return e;
    if (!h.empty()) e = cudaMemcpyAsync(d.d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st);
    return e;
}

int World::rebuild_topology() {
    const size_t nb = bodies.slots(), nc = cols.slots();
    topo_error = 0;
    topo_error_msg.clear();
    binfo.resize(nb, make_uint2(0u, (uint32_t)BODY_NO_COLLIDER)); bmg.resize(nb, make_float2(1.0f, 1.0f)); inertia.resize(nb, 1.0f);
    bworld.resize(nb, 0u);
    coff.resize(nc, make_float2(0.f, 0.f)); cconst.resize(nc, make_uint4(0u, 0u, 0u, 0u)); cparent.resize(nc, NO_SLOT);
    ccold.resize(nc, make_uint4(0u, 0u, 0u, NO_SLOT));

    // springs (slot order) -> per-body CSR in spring order
    std::vector<SpringParams> sp;
    for (auto& b : hb) { b.n_springs = 0; b.n_joints = 0; }
    for (uint32_t s = 0; s < springs.slots(); ++s) {
        if (!springs.alive[s]) continue;
        const HSpring& x = hs[s];
        if (!bodies.valid(x.a) || !bodies.valid(x.b)) { topo_error = BLOBS_ERR_DANGLING; topo_error_msg = "spring references a removed rigid body (zip_unwrap on None, springs.rs:26-29)"; continue; }
        if (h_slot(x.a) == h_slot(x.b)) { topo_error = BLOBS_ERR_SAME_BODY; topo_error_msg = "spring: get2_mut called with identical indices"; continue; }
        sp.push_back(SpringParams{h_slot(x.a), h_slot(x.b), x.rest, x.k, x.c});
        hb[h_slot(x.a)].n_springs++;
        hb[h_slot(x.b)].n_springs++;
    }
    n_springs_live = (uint32_t)sp.size();
    std::vector<uint32_t> v_sb_body, v_sb_off, v_sb_edge;
    {
        std::vector<int32_t> idx(nb, -1);
        for (uint32_t b = 0; b < nb; ++b)
            if (bodies.alive[b] && hb[b].n_springs) { idx[b] = (int32_t)v_sb_body.size(); v_sb_body.push_back(b); }
        v_sb_off.assign(v_sb_body.size() + 1, 0);
        for (size_t i = 0; i < v_sb_body.size(); ++i) v_sb_off[i + 1] = v_sb_off[i] + hb[v_sb_body[i]].n_springs;
        v_sb_edge.resize(v_sb_off.back());
        std::vector<uint32_t> fill(v_sb_off.begin(), v_sb_off.end() - 1);
        for (uint32_t i = 0; i < sp.size(); ++i) {
            v_sb_edge[fill[idx[sp[i].a]]++] = (i << 1) | 0u;
            v_sb_edge[fill[idx[sp[i].b]]++] = (i << 1) | 1u;
        }
    }
    n_sb = (uint32_t)v_sb_body.size();

    // joints (slot order) -> islands via union-find, joints of an island stay in slot order
    std::vector<JointParams> jp;
    std::vector<uint32_t> v_isl_off{0}, v_isl_joint;
    {
        std::vector<uint32_t> uf(nb);
        std::iota(uf.begin(), uf.end(), 0u);
        auto find = [&](uint32_t x) { while (uf[x] != x) { uf[x] = uf[uf[x]]; x = uf[x]; } return x; };
        for (uint32_t s = 0; s < joints.slots(); ++s) {
            if (!joints.alive[s]) continue;
            const HJoint& x = hj[s];
            if (!bodies.valid(x.a) || !bodies.valid(x.b)) { topo_error = BLOBS_ERR_DANGLING; topo_error_msg = "joint references a removed rigid body (unwrap on None, physics.rs:427-432)"; continue; }
            const uint32_t a = h_slot(x.a), b = h_slot(x.b);
            if (!(bmg.h[a].x > 0.0f) || !(bmg.h[b].x > 0.0f)) { topo_error = BLOBS_ERR_MASS; topo_error_msg = "assertion failed: calculated_mass > 0.0 (physics.rs:447-448)"; }
            jp.push_back(JointParams{a, b, x.aa.x, x.aa.y, x.ab.x, x.ab.y, x.distance, x.target});
            hb[a].n_joints++;
            hb[b].n_joints++;
            uf[find(a)] = find(b);
        }
        std::vector<int32_t> isl_of(nb, -1);
        std::vector<uint32_t> cnt;
        for (const auto& j : jp) {
            const uint32_t r = find(j.a);
            if (isl_of[r] < 0) { isl_of[r] = (int32_t)cnt.size(); cnt.push_back(0); }
            cnt[isl_of[r]]++;
        }
        v_isl_off.assign(cnt.size() + 1, 0);
        for (size_t i = 0; i < cnt.size(); ++i) v_isl_off[i + 1] = v_isl_off[i] + cnt[i];
        v_isl_joint.resize(jp.size());
        std::vector<uint32_t> fill(v_isl_off.begin(), v_isl_off.end() - 1);
        for (uint32_t i = 0; i < jp.size(); ++i) v_isl_joint[fill[isl_of[find(jp[i].a)]]++] = i;
        n_islands = (uint32_t)cnt.size();
    }
    n_joints_live = (uint32_t)jp.size();
    // joint records with island-local body indices (bodies numbered in order of first appearance), interleaved per CTA of
    // JOINT_THREADS islands: record e of island i at ((i / T) * max_j + e) * T + (i % T), two float4 each
    std::vector<uint32_t> v_isl_boff{0}, v_isl_body;
    std::vector<float4> jinter;
    isl_max_bodies = 0;
    isl_max_joints = 0;
    for (uint32_t i = 0; i < n_islands; ++i) isl_max_joints = std::max(isl_max_joints, v_isl_off[i + 1] - v_isl_off[i]);
    const size_t n_jblocks = (n_islands + JOINT_THREADS - 1) / JOINT_THREADS;
    const bool inter_ok = n_islands > 0 && (double)n_jblocks * isl_max_joints * JOINT_THREADS <= 4.0 * (double)jp.size() + 65536.0;
    if (inter_ok) jinter.assign(2 * n_jblocks * isl_max_joints * JOINT_THREADS, make_float4(0.f, 0.f, 0.f, 0.f));
    {
        std::vector<int32_t> local(nb, -1);
        for (uint32_t i = 0; i < n_islands; ++i) {
            const size_t first = v_isl_body.size();
            for (uint32_t e = v_isl_off[i]; e < v_isl_off[i + 1]; ++e) {
                JointParams j = jp[v_isl_joint[e]];
                for (uint32_t* s2 : {&j.a, &j.b}) {
                    if (local[*s2] < 0) { local[*s2] = (int32_t)(v_isl_body.size() - first); v_isl_body.push_back(*s2); }
                    *s2 = (uint32_t)local[*s2];
                }
                if (inter_ok) {
                    const size_t idx = 2 * (((size_t)(i / JOINT_THREADS) * isl_max_joints + (e - v_isl_off[i])) * JOINT_THREADS + (i % JOINT_THREADS));
                    float fa, fb;
                    std::memcpy(&fa, &j.a, 4);
                    std::memcpy(&fb, &j.b, 4);
                    jinter[idx] = make_float4(fa, fb, j.aax, j.aay);
                    jinter[idx + 1] = make_float4(j.abx, j.aby, j.distance, j.target);
                }
            }
            for (size_t q = first; q < v_isl_body.size(); ++q) local[v_isl_body[q]] = -1;
            isl_max_bodies = std::max<uint32_t>(isl_max_bodies, (uint32_t)(v_isl_body.size() - first));
            v_isl_boff.push_back((uint32_t)v_isl_body.size());
        }
    }
    joints_smem_ok = inter_ok && isl_max_bodies <= (uint32_t)JOINT_SMEM_MAX;

    // colliders: parent resolution
    r_max = 0.f;
    n_active_cols = 0;
    for (uint32_t c = 0; c < nc; ++c) {
        uint32_t f = 0, p = NO_SLOT;
        if (cols.alive[c]) {
            const HCollider& x = hc[c];
            if (x.parent != 0 && bodies.valid(x.parent)) {
                f |= CF_ACTIVE;
                p = h_slot(x.parent);
                r_max = std::max(r_max, x.desc.radius);
                n_active_cols++;
            }
            if (x.desc.is_sensor) f |= CF_SENSOR;
            if (x.desc.offset.translation.x != 0.0f || x.desc.offset.translation.y != 0.0f) f |= CF_OFFSET;
            // default cold half? (sole collider of its body, mass exactly 4r, groups ALL, not a sensor) — otherwise the narrowphase
            // must fetch ccold[]. Event recording needs the partner's real parent slot, so it flags everything.
            // (a collider with an offset is flagged too: k_tile takes "not flagged" to mean "snapshot == body position")
            bool dflt = p != NO_SLOT && !x.desc.is_sensor && x.desc.memberships == 0xffffffffu && x.desc.filter == 0xffffffffu &&
                        rec_mode != BLOBS_RECORD_EVENTS && !(f & CF_OFFSET);
            if (dflt) {
                const HBody& pb = hb[p];
                size_t live = 0;
                for (uint32_t cc2 : pb.cols) live += (cols.alive[cc2] && hc[cc2].parent == x.parent) ? 1 : 0;
                const float m4 = 4.0f * x.desc.radius, mb = bmg.h[p].x;
                dflt = live == 1 && std::memcmp(&m4, &mb, 4) == 0;
            }
            if (!dflt) f |= CF_COLD;
        }
        {
            uint4 e = cconst.h[c];
            e.y = f;
            cconst.set(c, e);
        }
        cparent.set(c, p);
        {
            uint32_t mbits = 0;
            if (p != NO_SLOT) { const float m = bmg.h[p].x; std::memcpy(&mbits, &m, 4); }
            const uint4 e = cconst.h[c];
            ccold.set(c, make_uint4(mbits, e.z, e.w, p));
        }
    }

    // bodies
    std::vector<uint32_t> v_mb_body, v_mb_off{0}, v_mb_cols;
    any_dynamic = false;
    std::vector<uint8_t> world_has_first(n_worlds, 0);
    n_simple = 0;
    n_loose = 0;
    for (uint32_t b = 0; b < nb; ++b) {
        uint32_t f = 0;
        int32_t bc = BODY_NO_COLLIDER;
        if (bodies.alive[b]) {
            HBody& x = hb[b];
            f |= BF_ALIVE;
            // physics.rs:327-339: the dt/old_dt ratio goes to the first NON-STATIC body in arena order (kinematic bodies count as
            // non-static), and old_dt is only overwritten when such a body exists
            if (x.type == BLOBS_BODY_STATIC) {
                f |= BF_STATIC;
            } else {
                any_dynamic = true;
                if (!world_has_first[x.world]) { world_has_first[x.world] = 1; f |= BF_FIRST_DYN; }
            }
            if (x.type == BLOBS_BODY_KINEMATIC_POSITION || x.type == BLOBS_BODY_KINEMATIC_VELOCITY) f |= BF_KINEMATIC;
            if (x.n_springs) f |= BF_SPRINGS;
            if (x.n_joints) f |= BF_JOINTED | BF_ROT;
            if (x.rot_active) f |= BF_ROT;
            // distinct live colliders parented to this body, ascending slot
            auto& cs = x.cols;
            cs.erase(std::remove_if(cs.begin(), cs.end(), [&](uint32_t c) { return !cols.alive[c] || hc[c].parent != bodies.handle_at(b); }), cs.end());
            if (cs.size() == 1) { bc = (int32_t)cs[0]; n_simple++; }
            else if (cs.size() > 1) {
                std::sort(cs.begin(), cs.end());
                bc = -(int32_t)v_mb_body.size() - 2;
                v_mb_body.push_back(b);
                v_mb_cols.insert(v_mb_cols.end(), cs.begin(), cs.end());
                v_mb_off.push_back((uint32_t)v_mb_cols.size());
            } else {   // no collider: only kernels with one thread per body SLOT reach it (k_main, k_integrate)
                n_simple++;
                if (!x.n_joints) { f |= BF_LOOSE; n_loose++; }   // (a jointed one is advanced by k_integrate(BF_JOINTED) anyway)
            }
        }
        binfo.set(b, make_uint2(f, (uint32_t)bc));
    }
    n_multi = (uint32_t)v_mb_body.size();

    CU(upload(mb_body, v_mb_body, stream)); CU(upload(mb_off, v_mb_off, stream)); CU(upload(mb_cols, v_mb_cols, stream));
    CU(upload(sb_body, v_sb_body, stream)); CU(upload(sb_off, v_sb_off, stream)); CU(upload(sb_edge, v_sb_edge, stream));
    CU(upload(isl_off, v_isl_off, stream)); CU(upload(isl_joint, v_isl_joint, stream));
    CU(upload(d_springs, sp, stream)); CU(upload(d_joints, jp, stream));
    CU(upload(d_joints_inter, jinter, stream)); CU(upload(isl_boff, v_isl_boff, stream)); CU(upload(isl_body, v_isl_body, stream));
    CU(cudaStreamSynchronize(stream));  // the staging vectors above are locals
    topo_dirty = false;
    bp_dirty = true;  // record words (flags, parents, masses) may have changed
    return BLOBS_OK;
}

int World::flush() {
    {
        int rc0 = ensure_capacity();
        if (rc0) return rc0;
    }
    if (topo_dirty) {
        int rc = rebuild_topology();
        if (rc) return rc;
    }
    CU(inertia.flush(stream)); CU(binfo.flush(stream)); CU(bmg.flush(stream)); CU(bworld.flush(stream));
    CU(coff.flush(stream)); CU(cconst.flush(stream)); CU(cparent.flush(stream)); CU(ccold.flush(stream));
    if (con_dirty) {
        std::vector<float4> kc(con_pos.size());
        for (size_t i = 0; i < kc.size(); ++i) kc[i] = make_float4(con_pos[i].x, con_pos[i].y, con_r[i], 0.f);
        CU(d_constraints.ensure(std::max<size_t>(kc.size(), 1), stream));
        if (!kc.empty()) CU(cudaMemcpyAsync(d_constraints.d, kc.data(), kc.size() * sizeof(float4), cudaMemcpyHostToDevice, stream));
        CU(cudaStreamSynchronize(stream));
        con_dirty = false;
    }
    int rc = flush_writes();
    if (rc) return rc;
    CU(cudaStreamSynchronize(stream));  // Mirrored uploads read pageable host vectors
    if (bp_dirty) {
        rc = rebuild_broadphase();
        if (rc) return rc;
    }
    return BLOBS_OK;
}

// ---------------------------------------------------------------------------------------------- broadphase
int World::choose_grid(bool) {
    const size_t nc = cols.slots();
    // list pipeline: the rebuild collects every collider within r_a + r_b + skin, so the search reach (and the cell that keeps it
    // inside a 3x3 neighbourhood) grows by the skin
    nl_on = (list_mode == 1 || (list_mode == 2 && nl_grid_hold == 0)) && (!strip_on || nls_ready);
    nl_skin = nl_on ? skin_frac * r_max : 0.f;
    float cs = bp_cell_override > 0.f ? bp_cell_override : (r_max > 0.f ? 2.0f * r_max + nl_skin : 1.0f);
    if (!(cs > 0.f) || !std::isfinite(cs)) cs = 1.0f;
    DeviceStats init{};
    init.bb_min_x = init.bb_min_y = INT32_MAX;
    init.bb_max_x = init.bb_max_y = INT32_MIN;
    *h_stats = init;
    CU(cudaMemcpyAsync(d_stats, h_stats, sizeof(DeviceStats), cudaMemcpyHostToDevice, stream));
    if (nc) {
        BLOBS_LAUNCH(std::min(cdiv(nc, 256), 1184u), 256, 0, stream, k_bbox)(col_arrays(), cs, (uint32_t)nc, d_stats, strip_on ? d_cowned.d : nullptr);
        launches++;
        CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(h_stats, d_stats, sizeof(DeviceStats), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    long long ex = 1, ey = 1;
    if (h_stats->bb_min_x <= h_stats->bb_max_x) {
        ex = (long long)h_stats->bb_max_x - h_stats->bb_min_x + 1;
        ey = (long long)h_stats->bb_max_y - h_stats->bb_min_y + 1;
    }
    // a little slack so slow drift does not alias immediately; aliasing is harmless for correctness (toroidal table)
    double W = (double)ex + std::max(4.0, ex / 16.0) + (strip_on ? 6.0 : 0.0), H = (double)ey + std::max(4.0, ey / 16.0);
    // table budget: ~4 cells per collider (per batched world), at most 2^30 entries overall
    double cap = std::max<double>(n_worlds > 1 ? 64.0 : 4096.0, 4.0 * (double)std::max<uint32_t>(n_active_cols, 1) / (double)n_worlds);
    cap = std::min(cap, 1073741824.0 / (double)n_worlds);
    if (W * H > cap) {
        const double sc = std::sqrt(cap / (W * H));
        W = std::max(1.0, std::floor(W * sc));
        H = std::max(1.0, std::floor(H * sc));
    }
    grid.W = (uint32_t)W;
    grid.H = (uint32_t)H;
    grid.ncells = grid.W * grid.H;
    grid.n_worlds = n_worlds;
    grid.cell = cs;
    grid.inv_cell = 1.0f / cs;
    grid.rmax = nl_on ? (r_max + nl_skin) * 1.000001f : r_max;
    strip.rmax = grid.rmax;   // ghost selection reaches as far as the contact search does
    grid.MW = ~0ull / grid.W + 1ull;
    grid.MH = ~0ull / grid.H + 1ull;
    bb[0] = h_stats->bb_min_x; bb[1] = h_stats->bb_min_y; bb[2] = h_stats->bb_max_x; bb[3] = h_stats->bb_max_y;
    return BLOBS_OK;
}

int World::rebuild_broadphase() {
    int rc = choose_grid(true);
    if (rc) return rc;
    const size_t nc = cols.slots();
    const size_t tn = table_entries();
    CU(tab_a.ensure(tn + SCAN_ITEMS, stream)); CU(tab_b.ensure(tn + SCAN_ITEMS, stream));
    const size_t nrec = nc + 1 + (strip_on ? 2 * (size_t)strip.gcap : 0);  // +1: the scan may re-read index == #records
    CU(hot_a.ensure(nrec, stream)); CU(hot_b.ensure(nrec, stream));
    const unsigned ntiles = cdiv(tn, SCAN_TILE);
    CU(tile_a.ensure(ntiles, stream)); CU(tile_b.ensure(ntiles, stream));
    CU(cudaMemsetAsync(tab_a.d, 0, tn * sizeof(uint32_t), stream));
    CU(cudaMemsetAsync(tab_b.d, 0, tn * sizeof(uint32_t), stream));
    CU(cudaMemsetAsync(tile_a.d, 0, tile_a.cap * sizeof(uint32_t), stream));
    CU(cudaMemsetAsync(tile_b.d, 0, tile_b.cap * sizeof(uint32_t), stream));
    uint32_t* tab_next = cur_is_a ? tab_b.d : tab_a.d;
    uint32_t* tab_cur = cur_is_a ? tab_a.d : tab_b.d;
    uint32_t* tile_next = cur_is_a ? tile_b.d : tile_a.d;
    uint32_t* tile_cur = cur_is_a ? tile_a.d : tile_b.d;
    float4* hot_next = cur_is_a ? hot_b.d : hot_a.d;
    if (nl_on) {
        // list pipeline: the tables stay empty here; the first substep finds `force` set and runs the rebuild chain on the device
        const size_t ncap = std::max<size_t>(cabs.cap, 1);
        CU(snap_a.ensure(ncap, stream)); CU(snap_b.ensure(ncap, stream)); CU(nl_hdr.ensure(ncap, stream));
        nl_stride = (uint32_t)nl_hdr.cap;
        CU(nl_idx.ensure((size_t)NL_CAP * nl_stride, stream));
        NlCtl init{};
        init.force = 1u;
        init.rebuilds = h_nlctl->rebuilds;
        init.substeps = h_nlctl->substeps;
        init.pub_seq = h_nlctl->pub_seq;   // (strips) sequence numbers stay monotonic: the flag slots still hold the old ones
        *h_nlctl = init;
        CU(cudaMemcpyAsync(d_nlctl, h_nlctl, sizeof(NlCtl), cudaMemcpyHostToDevice, stream));
        CU(cudaStreamSynchronize(stream));
        nl_force_pending = false;
        bp_dirty = false;
        return BLOBS_OK;
    }
    if (nc) {
        BLOBS_LAUNCH(cdiv(nc, 256), 256, 0, stream, k_count)(grid, col_arrays(), bworld.d.d, tab_next, tile_next, (uint32_t)nc, strip_on ? d_cowned.d : nullptr);
        launches++;
    }
    {
        int rc2 = strip_build_tail(tab_next, tab_cur, tile_next, tile_cur, hot_next, false);
        if (rc2) return rc2;
    }
    CU(cudaGetLastError());
    cur_is_a = !cur_is_a;
    bp_dirty = false;
    return BLOBS_OK;
}

// ---------------------------------------------------------------------------------------------- neighbour lists
NlView World::nl_view() {
    NlView L{};
    L.snap_cur = cur_is_a ? snap_a.d : snap_b.d;
    L.snap_next = cur_is_a ? snap_b.d : snap_a.d;
    L.hdr = nl_hdr.d;
    L.idx = nl_idx.d;
    L.stride = nl_stride;
    L.skin = nl_skin;
    // while collisions are disabled nothing reads the lists: no rebuilds (re-enabling forces one, see step())
    L.lim = collisions_enabled ? 0.45f * nl_skin : INFINITY;
    L.ctl = d_nlctl;
    L.tab[0] = tab_a.d; L.tab[1] = tab_b.d;
    L.tile[0] = tile_a.d; L.tile[1] = tile_b.d;
    L.hot = hot_a.d;
    if (strip_on) {
        const size_t off = cur_is_a ? nls_snap_cap * sizeof(float4) : 0;   // the neighbours' snap_next: buffers flip in lock step on every rank
        L.peer_next[0] = nls_peer_snap[0] ? reinterpret_cast<float4*>(nls_peer_snap[0] + off) : nullptr;
        L.peer_next[1] = nls_peer_snap[1] ? reinterpret_cast<float4*>(nls_peer_snap[1] + off) : nullptr;
        L.olist = olist.d;
        L.ocount = d_ocount;
    }
    return L;
}

NlStripDev World::nls_dev() {
    NlStripDev X{};
    X.rank = s_rank;
    X.nranks = s_nranks;
    X.mine = nls_flags;
    for (int r = 0; r < NL_MAX_RANKS; ++r) X.peer[r] = nls_peer_flags[r];
    X.recv_block = p2p_block;
    X.peer_block[0] = p2p_peer[0];
    X.peer_block[1] = p2p_peer[1];
    X.stride = p2p_stride;
    X.xseq = d_push_done + 1;
    X.push_done = d_push_done;
    return X;
}

// First launches of every substep in list mode: the decision kernel, then the four rebuild kernels, which return at once unless
// the decision was "rebuild" (the decision lives on the device, so that captured steps can be replayed).
int World::nl_rebuild_chain(bool timed_launch, bool decide) {
    const NlView L = nl_view();
    const ColliderArrays C = col_arrays();
    const BodyArrays B = body_arrays();
    const uint32_t nc = (uint32_t)cols.slots();
    const size_t tn = table_entries();
    auto run = [&](KClass k, auto&& f) -> int {
        if (timed_launch) return timed(k, f);
        f();
        launches++;
        CU(cudaGetLastError());
        return BLOBS_OK;
    };
    int rc = BLOBS_OK;
    // Inside a graph capture the rebuild kernels go into the body of an IF node whose condition the deciding kernel sets on the
    // device (cudaGraphSetConditional); otherwise they are launched unconditionally and return at once when NlCtl::need is 0.
    unsigned long long h_cur = 0ull;
    bool use_cond = false;
#ifndef BLOBS_EMU
    cudaGraph_t cap_graph = nullptr;
    use_cond = capturing && cond_nodes && !profiling && timed_launch && nc != 0;
    if (use_cond) {
        cudaStreamCaptureStatus cs;
        CU(cudaStreamGetCaptureInfo(stream, &cs, nullptr, &cap_graph, nullptr, nullptr));
        if (nl_cond_pending) {   // created by the previous substep, whose k_step sets it
            h_cur = nl_cond_pending;
            nl_cond_pending = 0ull;
        } else {
            cudaGraphConditionalHandle h;
            CU(cudaGraphConditionalHandleCreate(&h, cap_graph, 0, cudaGraphCondAssignDefault));
            h_cur = (unsigned long long)h;
        }
    }
#endif
    const NlStripDev X = nls_dev();
    const uint8_t* cown = strip_on ? d_cowned.d : nullptr;
    // strips: the rebuild kernels enumerate the owned-body list (work proportional to the strip, not to the world's slot count)
    NlEnum E{};
    E.n = nc;
    if (strip_on) { E.olist = olist.d; E.ocount = d_ocount; E.binfo = binfo.d.d; E.n = std::max<uint32_t>(olaunch_dim, 1); }
    if (strip_on) {
        // Strips: the decision combines every rank's numbers of the previous substep (and so waits for their ghost records)
        rc = run(KC_DECIDE, [&] { BLOBS_LAUNCH(1, 32, 0, stream, k_nls_decide)(d_nlctl, X, L.lim, 0.25f * nl_skin, msg[0], msg[1], timed_launch ? 1u : 0u, d_stats, h_cur); });
    } else if (decide) {
        rc = run(KC_DECIDE, [&] { BLOBS_LAUNCH(1, 32, 0, stream, k_nl_decide)(d_nlctl, L.lim, timed_launch ? 1u : 0u, h_cur); });
    }
    if (rc) return rc;
    if (!nc) return BLOBS_OK;
    const uint32_t ne = E.n;
    auto rebuild = [&]() -> int {
        int r;
        if (strip_on) {   // a rebuild re-selects ghosts and hands migrants over with the message exchange of the grid pipeline
            r = run(KC_PACK, [&] { BLOBS_LAUNCH(std::min(cdiv(ne, 256), NL_GATED_CTAS), 256, 0, stream, k_nls_pack)(B, C, strip, E, msg[0], msg[1], d_nlctl); });
            if (r) return r;
            r = run(KC_NCCL, [&] { BLOBS_LAUNCH(STRIP_PUSH_CTAS, 256, 0, stream, k_nls_push)(strip, msg[0], msg[1], X, d_nlctl, d_stats); });
            if (r) return r;
        }
        r = run(KC_SCAN, [&] { BLOBS_LAUNCH(std::min(cdiv(ne, 256), NL_GATED_CTAS), 256, 0, stream, k_nl_count)(grid, C, bworld.d.d, L, E); });
        if (r) return r;
        if (strip_on) {
            r = run(KC_GHOST, [&] { BLOBS_LAUNCH(cdiv(2 * (size_t)strip.gcap, 256), 256, 0, stream, k_nls_bin_ghosts)(grid, strip, X, L, gcell.d, d_stats); });
            if (r) return r;
        }
        r = run(KC_SCAN, [&] { BLOBS_LAUNCH(cdiv(tn, SCAN_TILE), SCAN_THREADS, 0, stream, k_nl_scan)(L, (uint32_t)tn); });
        if (r) return r;
        r = run(KC_SCATTER, [&] { BLOBS_LAUNCH(std::min(cdiv(ne, 256), NL_GATED_CTAS), 256, 0, stream, k_nl_scatter)(C, L, E); });
        if (r) return r;
        if (strip_on) {
            r = run(KC_GHOST, [&] {
                BLOBS_LAUNCH(cdiv(2 * (size_t)strip.gcap + 4 * (size_t)strip.mcap, 256), 256, 0, stream, k_nls_finish)(B, C, strip, msg[0], msg[1], X, L, gcell.d, d_owned.d, d_cowned.d, olist.d, d_ocount,
                                                                                                             opos.d, (uint32_t)olist.cap, d_stats, cur_is_a ? snap_a.d : snap_b.d);
            });
            if (r) return r;
        }
        return run(KC_NLBUILD, [&] {
            BLOBS_LAUNCH(std::min(cdiv(ne, NL_BUILD_THREADS), 2 * NL_GATED_CTAS), NL_BUILD_THREADS, 0, stream, k_nl_build)(grid, C, bworld.d.d, L, cur_is_a ? snap_a.d : snap_b.d, E, cown, strip);
        });
    };
#ifndef BLOBS_EMU
    if (use_cond) {
        cudaStreamCaptureStatus cs;
        const cudaGraphNode_t* deps = nullptr;
        size_t nd = 0;
        CU(cudaStreamGetCaptureInfo(stream, &cs, nullptr, &cap_graph, &deps, &nd));
        cudaGraphNodeParams np = {};   // (a union with non-trivial members: value-initialised, then the fields that matter)
        std::memset(static_cast<void*>(&np), 0, sizeof(np));
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = (cudaGraphConditionalHandle)h_cur;
        np.conditional.type = cudaGraphCondTypeIf;
        np.conditional.size = 1;
        cudaGraphNode_t cnode = nullptr;
        CU(cudaGraphAddNode(&cnode, cap_graph, deps, nd, &np));
        cudaGraph_t body = np.conditional.phGraph_out[0];
        CU(cudaStreamUpdateCaptureDependencies(stream, &cnode, 1, cudaStreamSetCaptureDependencies));
        // the body: the very same (still self-gating) rebuild kernels, captured into the IF node's graph through a second stream
        if (!s_body) CU(cudaStreamCreateWithFlags(&s_body, cudaStreamNonBlocking));
        CU(cudaStreamBeginCaptureToGraph(s_body, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        std::swap(stream, s_body);
        const uint64_t l0 = launches;
        rc = rebuild();
        cond_launches += launches - l0;
        launches = l0;   // counted per executed rebuild (finish_stats), not per replay
        std::swap(stream, s_body);
        cudaGraph_t out = nullptr;
        cudaError_t e = cudaStreamEndCapture(s_body, &out);
        if (rc) return rc;
        if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture (IF-node body)");
        cond_nodes_built++;
        return BLOBS_OK;
    }
#endif
    return rebuild();
}

// Outside a step (scene queries): bring the cell grid up to date with the current snapshots, and learn which table holds it.
int World::nl_rebuild_now() {
    const unsigned int one = 1u;
    CU(cudaMemcpyAsync(&d_nlctl->force, &one, sizeof(one), cudaMemcpyHostToDevice, stream));
    int rc = nl_rebuild_chain(false, true);
    if (rc) return rc;
    CU(cudaMemcpyAsync(h_nlctl, d_nlctl, sizeof(NlCtl), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    return BLOBS_OK;
}

// ---------------------------------------------------------------------------------------------- stepping
// profiler range per kernel class: the reference's tracy span names where the kernel replaces a spanned function
// (physics.rs:242 "brute_force_collisions", physics.rs:324 "update positions"), descriptive names otherwise
static const char* const kclass_span[KC_COUNT] = {"brute_force_collisions", "broadphase scan", "broadphase scatter", "springs", "solve_fixed_joints",
                                                  "update positions", "other", "strip pack", "strip ghosts", "strip exchange", "crowded contacts", "neighbour lists", "list decision"};

template <class F>
int World::timed(KClass k, F&& f) {
    Span span(kclass_span[k]);
    EvPair* ep = nullptr;
    EvPair cap{};
    const bool timed_here = profiling && (!profile_main_only || k == KC_MAIN || k == KC_CROWDED);
    if (timed_here && capturing) {
        // inside a graph capture: timing events become external event-record nodes, re-recorded by every replay
        CU(cudaEventCreate(&cap.a));
        CU(cudaEventCreate(&cap.b));
        cap.k = k;
        CU(cudaEventRecordWithFlags(cap.a, stream, cudaEventRecordExternal));
    } else if (timed_here) {
        if (ev_used == ev_pool.size()) {
            EvPair p{};
            CU(cudaEventCreate(&p.a));
            CU(cudaEventCreate(&p.b));
            ev_pool.push_back(p);
        }
        ep = &ev_pool[ev_used++];
        ep->k = k;
        CU(cudaEventRecord(ep->a, stream));
    }
    f();
    launches++;
    if (ep) CU(cudaEventRecord(ep->b, stream));
    if (timed_here && capturing) {
        CU(cudaEventRecordWithFlags(cap.b, stream, cudaEventRecordExternal));
        cap_evs->push_back(cap);
    }
    CU(cudaGetLastError());
    return BLOBS_OK;
}

int World::collect_profile() {
    for (size_t i = 0; i < ev_used; ++i) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, ev_pool[i].a, ev_pool[i].b));
        prof_ms[ev_pool[i].k] += ms;
        prof_launches[ev_pool[i].k]++;
    }
    ev_used = 0;
    for (GraphSlot* g : graphs_launched) {
        if (!g->profiled) continue;
        for (const EvPair& e : g->evs) {
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, e.a, e.b));
            prof_ms[e.k] += ms;
            prof_launches[e.k]++;
        }
    }
    graphs_launched.clear();
    return BLOBS_OK;
}

void World::destroy_graph(GraphSlot& g) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    for (auto& e : g.evs) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    g.evs.clear();
    g.exec = nullptr;
    g.key = 0;
}

// Everything that is baked into the kernel arguments of one Physics::integrate call. If any of it changes the graph is
// re-captured; body/collider DATA changes (staged writes, forces) do not invalidate it.
uint64_t World::step_key(uint32_t nsub, float delta, bool last) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) {
        const unsigned char* c = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 1099511628211ull; }
    };
#define MIXV(v) { auto t__ = (v); mix(&t__, sizeof(t__)); }
    MIXV(nsub) MIXV(delta) MIXV(last) MIXV(old_dt) MIXV(gx) MIXV(gy) MIXV(collisions_enabled) MIXV(joint_iterations) MIXV(contact_mode)
    MIXV(allow_fused) MIXV(joint_advance) MIXV(tune) MIXV(crowded_mode) MIXV(crowded_seen) MIXV(pool_mode) MIXV(pool_seen) MIXV(pool_min) MIXV(over_list.d) MIXV(cur_is_a) MIXV(any_dynamic) MIXV(rec_mode) MIXV(profiling) MIXV(profile_main_only)
    MIXV(bodies.slots()) MIXV(cols.slots()) MIXV(con_pos.size()) MIXV(n_multi) MIXV(n_sb) MIXV(n_islands) MIXV(n_joints_live) MIXV(isl_max_bodies)
    MIXV(isl_max_joints) MIXV(joints_smem_ok) MIXV(grid) MIXV(strip_on) MIXV(strip) MIXV(olaunch_dim) MIXV(n_loose) MIXV(n_active_cols)
    const BodyArrays B = body_arrays();
    const ColliderArrays C = col_arrays();
    mix(&B, sizeof(B));
    mix(&C, sizeof(C));
    MIXV(hot_a.d) MIXV(hot_b.d) MIXV(tab_a.d) MIXV(tab_b.d) MIXV(tile_a.d) MIXV(tile_b.d) MIXV(d_constraints.d) MIXV(mb_body.d) MIXV(mb_off.d) MIXV(mb_cols.d)
    MIXV(sb_body.d) MIXV(sb_off.d) MIXV(sb_edge.d) MIXV(d_springs.d) MIXV(isl_off.d) MIXV(isl_joint.d) MIXV(d_joints.d) MIXV(d_joints_inter.d)
    MIXV(nl_on) MIXV(nl_skin) MIXV(snap_a.d) MIXV(snap_b.d) MIXV(nl_hdr.d) MIXV(nl_idx.d) MIXV(nl_stride) MIXV(d_nlctl)
    MIXV(isl_boff.d) MIXV(isl_body.d) MIXV(olist.d) MIXV(opos.d) MIXV(d_owned.d) MIXV(d_cowned.d) MIXV(gcell.d) MIXV(msg[0]) MIXV(msg[1]) MIXV(msg[2]) MIXV(msg[3]) MIXV(p2p_on) MIXV(p2p_block) MIXV(p2p_peer[0]) MIXV(p2p_peer[1]) MIXV(nccl_exchanges & 1u)
#undef MIXV
    return h ? h : 1;
}

// One Physics::integrate call: replay the captured graph when nothing structural changed, else (re)capture it.
int World::run_step(uint32_t nsub, float delta, bool last, bool allow_graph) {
    Span span("integrate");   // physics.rs:92,398
    // measured on 2x B200: replaying a graph that contains the grouped ncclSend/ncclRecv is ~25 % SLOWER than plain launches,
    // so strip mode keeps plain launches
    // (with the peer-memory exchange the substep holds no NCCL call; replaying it is opt-in until measured: BLOBS_B200_STRIP_GRAPH=1)
    if (!graphs_on || !allow_graph || nsub == 0 || rec_mode != BLOBS_RECORD_OFF || (strip_on && !(p2p_on && (strip_graph || nl_on)))) return integrate(nsub, delta, last);
    const uint64_t key = step_key(nsub, delta, last);
    GraphSlot& gs = gslot[last ? 1 : 0];
    const float step_delta = delta / (float)nsub;
    if (gs.exec && gs.key == key) {
        CU(cudaGraphLaunch(gs.exec, stream));
        // host-side effects of integrate()
        if (any_dynamic) old_dt = step_delta;
        if (nsub & 1u) cur_is_a = !cur_is_a;
        if (strip_on) nccl_exchanges += nsub;   // one exchange per substep; its parity picks the receive buffers baked into the graph
        launches += gs.launches;
        cond_per_rebuild_live = gs.cond_per_rebuild;
        graph_replays++;
        graphs_launched.push_back(&gs);
        return BLOBS_OK;
    }
    destroy_graph(gs);
    const uint64_t l0 = launches;
    cond_launches = 0;
    cond_nodes_built = 0;
    capturing = true;
    cap_evs = &gs.evs;
    cudaError_t e = cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { capturing = false; return cuda_fail(e, "cudaStreamBeginCapture"); }
    const int rc = integrate(nsub, delta, last);
    cudaGraph_t graph = nullptr;
    e = cudaStreamEndCapture(stream, &graph);
    capturing = false;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture");
    e = cudaGraphInstantiate(&gs.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { gs.exec = nullptr; return cuda_fail(e, "cudaGraphInstantiate"); }
    gs.key = key;
    gs.launches = launches - l0;
    gs.cond_per_rebuild = cond_nodes_built ? cond_launches / cond_nodes_built : 0;
    gs.profiled = profiling;
    graph_captures++;
    cond_per_rebuild_live = gs.cond_per_rebuild;
    CU(cudaGraphLaunch(gs.exec, stream));
    graphs_launched.push_back(&gs);
    return BLOBS_OK;
}

int World::launch_substep(const SubstepParams& P_in) {
    Span span("substep");     // physics.rs:402
    SubstepParams P = P_in;
    // contact-list overflows: deferred to k_crowded (one warp per body) when such bodies are expected, else resolved inline
    // Automatic mode stays off in strip mode for now: on real GPUs the pooled / crowded kernels were validated on one device
    // only (the 2-GPU parity run is pending); forced (mode 1) they pass the 2-rank strip test of the host-compiled build
    // (tests/test_multi_gpu.py, shell scene: overflowing bodies on both sides of the strip edge).
    const bool auto_ok = !strip_on;
    const bool pooled = contact_mode == 0 && collisions_enabled && (pool_mode == 1 || (pool_mode == 2 && pool_seen && auto_ok));
    // (automatic mode: the pooled k_main hands its big-neighbourhood bodies to k_crowded, so the two come together)
    const bool lists = nl_on;
    const bool crowded = lists ? (collisions_enabled && (crowded_mode == 1 || (crowded_mode == 2 && crowded_seen)))
                               : (contact_mode == 0 && collisions_enabled && (crowded_mode == 1 || (crowded_mode == 2 && (crowded_seen || pooled) && auto_ok)));
    P.crowded = crowded ? 1u : 0u;
    P.over_parity = cur_is_a ? 0u : 1u;
    P.over_list = over_list.d;
    P.pool_min = pool_min;
    const BodyArrays B = body_arrays();
    const ColliderArrays C = col_arrays();
    const Constraints K = constraints_pod();
    Broadphase bp{};
    bp.hot = cur_is_a ? hot_a.d : hot_b.d;
    bp.tab = cur_is_a ? tab_a.d : tab_b.d;
    bp.tab_next = cur_is_a ? tab_b.d : tab_a.d;
    bp.tile_next = cur_is_a ? tile_b.d : tile_a.d;
    if (lists) {   // the kernels that walk the grid pick the table of the last rebuild on the device (resolve_grid)
        bp.nl = nl_view();
        bp.hot = bp.nl.hot;
        bp.tab = nullptr;
        bp.tab_next = bp.tile_next = nullptr;
    }
    uint32_t* tile_cur = cur_is_a ? tile_a.d : tile_b.d;
    uint32_t* tab_cur = cur_is_a ? tab_a.d : tab_b.d;
    float4* hot_next = cur_is_a ? hot_b.d : hot_a.d;
    Recording R;
    R.mode = (uint32_t)rec_mode;
    R.cap = (uint32_t)std::min<size_t>(rec_cap, 0xffffffffu);
    R.count = d_rec_count;
    R.pairs = rec_pairs.d;
    R.vels = rec_vels.d;
    // fused = bodies are advanced (verlet + snapshot + clamp + binning) by the kernel that last touches their position:
    // k_main for free bodies, k_joints_fused for jointed ones. Needs event recording off (events read pre-update velocities of
    // OTHER bodies) and, with joints, islands small enough for the shared-memory solver.
    const bool fused = (allow_fused && rec_mode != BLOBS_RECORD_EVENTS && (n_joints_live == 0 || joints_smem_ok)) || strip_on;
    const bool ordered = contact_mode == 0;
    last_fused = fused;
    const uint32_t nb = P.n_bodies;
    int rc;
    if (lists) {   // decide on the device whether the lists are still supersets of the contact set; rebuild them if not
        // k_step takes the decision for the NEXT substep itself when nothing else publishes snapshots after it
        P.nl_tail_decide = (fused && nb && !n_islands && !n_multi && !crowded && !strip_on) ? 1u : 0u;
        P.nl_tail_publish = (fused && nb && !n_islands && !n_multi && !crowded && strip_on && nls_tail_publish) ? 1u : 0u;
        rc = nl_rebuild_chain(true, !nl_prev_tail);
        if (rc) return rc;
        nl_prev_tail = P.nl_tail_decide != 0u;
        P.nl_cond_next = 0ull;
#ifndef BLOBS_EMU
        if (capturing && cond_nodes && !profiling && P.nl_tail_decide && nl_sub_i + 1 < nl_sub_n && cols.slots()) {
            // this substep's k_step decides for the next one: the IF node of the next substep needs its handle now
            cudaStreamCaptureStatus cs;
            cudaGraph_t cg = nullptr;
            CU(cudaStreamGetCaptureInfo(stream, &cs, nullptr, &cg, nullptr, nullptr));
            cudaGraphConditionalHandle h;
            CU(cudaGraphConditionalHandleCreate(&h, cg, 0, cudaGraphCondAssignDefault));
            P.nl_cond_next = nl_cond_pending = (unsigned long long)h;
        }
#endif
    }
    if (n_sb) {
        rc = timed(KC_SPRINGS, [&] { BLOBS_LAUNCH(cdiv(n_sb, 128), 128, 0, stream, k_springs)(P, B, sb_body.d, sb_off.d, sb_edge.d, d_springs.d, n_sb); });
        if (rc) return rc;
    }
    if (strip_on && !lists) {
        CU(cudaMemsetAsync(msg[0], 0, sizeof(StripHeader), stream));
        CU(cudaMemsetAsync(msg[1], 0, sizeof(StripHeader), stream));
    }
    if (nb && lists) {
        rc = timed(KC_MAIN, [&] {
            // BLOBS_PARAM_TUNE 1: 3 CTAs per SM (85 registers, no spills) instead of 4 (64 registers)
            // Register cap of k_step = CTAs of 256 threads per SM. The kernel is latency-bound (dependent 16-byte gathers), so occupancy
            // pays until the spills bite. Measured on B200 (profiles/r2_notes.md, driver window of config #2 / one strip rank at 2 M):
            // 4 CTAs (64 registers, no spills) 0.497 / 1.149 ms, 5 (48) 0.464, 6 (40) 0.469 / 1.081, 7-8 (32, ~230 B spilled) 0.496 / 1.125.
            // BLOBS_PARAM_TUNE overrides: 1 -> 3 CTAs, 2 -> 5, 3 -> 6, 4 -> 7, 5 -> 8, 6 -> 4.
#define BLOBS_LAUNCH_STEP(MINB)                                                                                                                            \
    do {                                                                                                                                                   \
        if (strip_on) BLOBS_LAUNCH(cdiv(std::max<uint32_t>(olaunch_dim, 1), 256), 256, 0, stream, k_step<true, MINB, true>)(P, grid, K, B, C, bp, R, d_stats, nls_dev()); \
        else BLOBS_LAUNCH(cdiv(nb, 256), 256, 0, stream, k_step<true, MINB, false>)(P, grid, K, B, C, bp, R, d_stats, NlStripDev{});                          \
    } while (0)
            if (fused || strip_on) {
                switch (tune) {
                    case 1: BLOBS_LAUNCH_STEP(3); break;
                    case 2: BLOBS_LAUNCH_STEP(5); break;
                    case 3: BLOBS_LAUNCH_STEP(6); break;
                    case 4: BLOBS_LAUNCH_STEP(7); break;
                    case 5: BLOBS_LAUNCH_STEP(8); break;
                    case 6: BLOBS_LAUNCH_STEP(4); break;
                    default:
                        if (strip_on) BLOBS_LAUNCH_STEP(6);
                        else BLOBS_LAUNCH_STEP(5);
                        break;
                }
            }
#undef BLOBS_LAUNCH_STEP
            else BLOBS_LAUNCH(cdiv(nb, 256), 256, 0, stream, k_step<false, 4, false>)(P, grid, K, B, C, bp, R, d_stats, NlStripDev{});
        });
        if (rc) return rc;
    } else if (nb) {
        rc = timed(KC_MAIN, [&] {
            const unsigned gdim = cdiv(strip_on ? std::max<uint32_t>(olaunch_dim, 1) : nb, 256);
            const StripView sv = strip_view();
#define BLOBS_LAUNCH_MAIN(F, O, PL) BLOBS_LAUNCH(gdim, 256, 0, stream, k_main<F, O, 4, 4, PL>)(P, grid, K, B, C, bp, R, d_stats, sv)
            if (pooled) {   // contact-rich state: warp-cooperative resolution (2 candidates per lane in flight; 4 measured the same)
                if (fused) BLOBS_LAUNCH(gdim, 256, 0, stream, k_main<true, true, 2, 4, true>)(P, grid, K, B, C, bp, R, d_stats, sv);
                else BLOBS_LAUNCH(gdim, 256, 0, stream, k_main<false, true, 2, 4, true>)(P, grid, K, B, C, bp, R, d_stats, sv);
            } else if (fused) {
                if (ordered) BLOBS_LAUNCH_MAIN(true, true, false);
                else BLOBS_LAUNCH_MAIN(true, false, false);
            } else {
                if (ordered) BLOBS_LAUNCH_MAIN(false, true, false);
                else BLOBS_LAUNCH_MAIN(false, false, false);
            }
#undef BLOBS_LAUNCH_MAIN
        });
        if (rc) return rc;
    }
    if (n_multi) {
        rc = timed(KC_MAIN, [&] {
            const unsigned g = cdiv(n_multi, 128);
            if (fused) {
                if (ordered) BLOBS_LAUNCH(g, 128, 0, stream, k_multi<true, true>)(P, grid, K, B, C, bp, R, d_stats, mb_body.d, mb_off.d, mb_cols.d, n_multi);
                else BLOBS_LAUNCH(g, 128, 0, stream, k_multi<true, false>)(P, grid, K, B, C, bp, R, d_stats, mb_body.d, mb_off.d, mb_cols.d, n_multi);
            } else {
                if (ordered) BLOBS_LAUNCH(g, 128, 0, stream, k_multi<false, true>)(P, grid, K, B, C, bp, R, d_stats, mb_body.d, mb_off.d, mb_cols.d, n_multi);
                else BLOBS_LAUNCH(g, 128, 0, stream, k_multi<false, false>)(P, grid, K, B, C, bp, R, d_stats, mb_body.d, mb_off.d, mb_cols.d, n_multi);
            }
        });
        if (rc) return rc;
    }
    if (crowded && nb) {
        rc = timed(KC_CROWDED, [&] {
            const StripView sv = lists ? StripView{} : strip_view();   // list pipeline: nothing is packed per substep (peer stores instead)
            if (fused) BLOBS_LAUNCH(CROWD_GRID, 32 * CROWD_WARPS, 0, stream, k_crowded<true>)(P, grid, K, B, C, bp, R, d_stats, sv, mb_body.d, mb_off.d, mb_cols.d);
            else BLOBS_LAUNCH(CROWD_GRID, 32 * CROWD_WARPS, 0, stream, k_crowded<false>)(P, grid, K, B, C, bp, R, d_stats, sv, mb_body.d, mb_off.d, mb_cols.d);
        });
        if (rc) return rc;
    }
    const size_t jsmem = (size_t)JOINT_THREADS * std::max<uint32_t>(isl_max_bodies, 1) * sizeof(float4);
    if (fused) {
        if (n_islands) {  // joint projection from shared memory, then a body-parallel (coalesced) verlet pass over the jointed bodies
            // joint_advance: the joint kernel also advances its bodies (no second pass that re-reads them)
            rc = timed(KC_JOINTS, [&] {
                BLOBS_LAUNCH(cdiv(n_islands, JOINT_THREADS), JOINT_THREADS, jsmem, stream, k_joints_fused)(P, B, isl_off.d, d_joints_inter.d, isl_max_joints, isl_boff.d, isl_body.d, n_islands,
                                                                                                  joint_iterations, d_stats, joint_advance ? 1u : 0u, grid, K, C, bp, mb_off.d, mb_cols.d);
            });
            if (rc) return rc;
            if (!joint_advance) {
                rc = timed(KC_INTEGRATE, [&] { BLOBS_LAUNCH(cdiv(nb, 256), 256, 0, stream, k_integrate)(P, grid, K, B, C, bp, d_stats, mb_off.d, mb_cols.d, (uint32_t)BF_JOINTED); });
                if (rc) return rc;
            }
        }
    } else {
        if (n_islands && joint_iterations) {
            if (joints_smem_ok) {
                rc = timed(KC_JOINTS, [&] {
                    BLOBS_LAUNCH(cdiv(n_islands, JOINT_THREADS), JOINT_THREADS, jsmem, stream, k_joints_fused)(P, B, isl_off.d, d_joints_inter.d, isl_max_joints, isl_boff.d, isl_body.d, n_islands,
                                                                                                      joint_iterations, d_stats, 0u, grid, K, C, bp, mb_off.d, mb_cols.d);
                });
            } else {
                rc = timed(KC_JOINTS, [&] { BLOBS_LAUNCH(cdiv(n_islands, 128), 128, 0, stream, k_joints)(P, B, isl_off.d, isl_joint.d, d_joints.d, n_islands, joint_iterations, d_stats); });
            }
            if (rc) return rc;
        }
        if (nb) {
            rc = timed(KC_INTEGRATE, [&] { BLOBS_LAUNCH(cdiv(nb, 256), 256, 0, stream, k_integrate)(P, grid, K, B, C, bp, d_stats, mb_off.d, mb_cols.d, 0u); });
            if (rc) return rc;
        }
    }
    if (!lists) {
        rc = strip_build_tail(bp.tab_next, tab_cur, bp.tile_next, tile_cur, hot_next, true);
        if (rc) return rc;
    } else if (strip_on && !P.nl_tail_publish) {   // end of the substep: this rank's displacement numbers + "my ghost records are out" go to every rank
        rc = timed(KC_GHOST, [&] { BLOBS_LAUNCH(1, 32, 0, stream, k_nls_publish)(d_nlctl, nls_dev()); });
        if (rc) return rc;
    }
    cur_is_a = !cur_is_a;   // list pipeline: swaps the slot-indexed snapshot buffers
    if (rec_mode && sub_recorded < d_sub_end.cap) {
        CU(cudaMemcpyAsync(d_sub_end.d + sub_recorded, d_rec_count, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream));
        sub_recorded++;
    }
    return BLOBS_OK;
}

// Physics::integrate (physics.rs:397-422)
int World::integrate(uint32_t nsub, float delta, bool last_of_call) {
    const float step_delta = delta / (float)nsub;
    nl_prev_tail = false;   // the first substep of a call always runs k_nl_decide (host requests are honoured there)
    nl_cond_pending = 0ull;
    for (uint32_t i = 0; i < nsub; ++i) {
        SubstepParams P;
        P.nl_tail_decide = 0u;
        P.nl_tail_publish = 0u;
        P.nl_cond_next = 0ull;
        P.acc_zero = (i > 0 && n_sb == 0) ? 1u : 0u;   // (strips: a migrant was advanced by its previous owner in the same substeps, so the same holds for it)
        nl_sub_i = i;
        nl_sub_n = nsub;
        P.dt = step_delta;
        P.ratio_first = step_delta / old_dt;        // physics.rs:338
        P.ratio_rest = step_delta / step_delta;     // every later body sees old_dt == dt (Q2)
        P.gx = gx;
        P.gy = gy;
        P.collisions_enabled = collisions_enabled ? 1u : 0u;
        P.n_bodies = (uint32_t)bodies.slots();
        P.n_colliders = (uint32_t)cols.slots();
        P.write_vel = ((last_of_call && i + 1 == nsub) || n_sb > 0 || rec_mode == BLOBS_RECORD_EVENTS) ? 1u : 0u;
        int rc = launch_substep(P);
        if (rc) return rc;
        if (any_dynamic) old_dt = step_delta;  // physics.rs:339
    }
    return BLOBS_OK;
}

int World::finish_stats(BlobsStepStats* out, uint32_t steps, uint32_t substeps_run) {
    const size_t nc = cols.slots();
    if (nc) {
        BLOBS_LAUNCH(std::min(cdiv(nc, 256), 1184u), 256, 0, stream, k_bbox)(col_arrays(), grid.cell, (uint32_t)nc, d_stats, strip_on ? d_cowned.d : nullptr);
        launches++;
    }
    CU(cudaEventRecord(ev_step1, stream));
    CU(cudaMemcpyAsync(h_stats, d_stats, sizeof(DeviceStats), cudaMemcpyDeviceToHost, stream));
    if (nl_on) CU(cudaMemcpyAsync(h_nlctl, d_nlctl, sizeof(NlCtl), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    CU(cudaGetLastError());
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev_step0, ev_step1);
    if (profiling) { int rc = collect_profile(); if (rc) return rc; }
    graphs_launched.clear();
    // perf_counter_inc("collisions", count) once per brute_force_collisions call (physics.rs:316); the per-substep counts of this
    // API call arrive here as one sum. In strip mode every rank adds the pairs it counted (its share of the world's).
    if (substeps_run && collisions_enabled) PerfCounters::global().inc("collisions", h_stats->collisions);
    last_max_ghosts = std::max(last_max_ghosts, h_stats->max_ghosts);
    last_max_migrants = std::max(last_max_migrants, h_stats->max_migrants);
    if (out) {
        out->collisions = h_stats->collisions;
        out->coincident_pairs = h_stats->coincident;
        out->events_dropped = h_stats->rec_dropped;
        out->nan_detected = h_stats->nan_flag;
        out->steps_run = steps;
        out->substeps_run = substeps_run;
        out->list_overflow = h_stats->list_overflow;
        out->gpu_ms = ms;
    }
    // auto mode: keep k_crowded in the pipeline for a while after the last overflow (a flip re-captures the CUDA graph)
    if (h_stats->list_overflow != 0) crowded_hold = 64;
    else if (crowded_hold > 0) crowded_hold--;
    crowded_seen = crowded_hold > 0;
    if (substeps_run) snap_epoch++;
    // automatic broadphase choice (BLOBS_PARAM_LIST 2): neighbour lists pay while they survive several substeps; when the scene is
    // so agitated that they are rebuilt (almost) every substep, sorting straight into cells every substep is the cheaper way to the
    // same contact set. Lists are tried again after a hold that doubles each time they turn out to be still too short-lived.
    // (strips: every rank sees the same rebuild counts - the decision is collective - so all ranks switch in the same call)
    if (list_mode == 2 && (!strip_on || nls_ready) && substeps_run) {
        if (nl_on) {
            const double dr = (double)(h_nlctl->rebuilds - nl_seen_rebuilds), ds = (double)(h_nlctl->substeps - nl_seen_substeps);
            if (ds > 0 && dr > 0.5 * ds + 1.0) {
                nl_grid_hold = nl_next_hold;
                nl_next_hold = std::min(nl_next_hold * 2, 1024);
                bp_dirty = true;
            } else if (ds > 0 && dr < 0.25 * ds) {
                nl_next_hold = 32;
            }
        } else if (nl_grid_hold > 0 && --nl_grid_hold == 0) {
            bp_dirty = true;
        }
    }
    if (nl_on && cond_per_rebuild_live) launches += (h_nlctl->rebuilds - nl_seen_rebuilds) * cond_per_rebuild_live;   // IF-node bodies that ran
    cond_per_rebuild_live = 0;
    nl_seen_rebuilds = h_nlctl->rebuilds;
    nl_seen_substeps = h_nlctl->substeps;
    // same for the warp-pooled k_main: worth it from ~0.25 contact pairs per body-substep
    if (substeps_run && (double)h_stats->collisions >= 0.25 * (double)substeps_run * (double)std::max<size_t>(bodies.slots(), 1)) pool_hold = 64;
    else if (pool_hold > 0) pool_hold--;
    pool_seen = pool_hold > 0;
    // table re-dimensioning: only when the snapshot outgrew (aliasing) or vastly undershoots the table
    if (h_stats->bb_min_x <= h_stats->bb_max_x) {
        const long long ex = (long long)h_stats->bb_max_x - h_stats->bb_min_x + 1, ey = (long long)h_stats->bb_max_y - h_stats->bb_min_y + 1;
        const double cap = std::max<double>(n_worlds > 1 ? 64.0 : 4096.0, 4.0 * (double)std::max<uint32_t>(n_active_cols, 1) / (double)n_worlds);
        const bool aliased = (ex > grid.W || ey > grid.H) && (double)grid.ncells < 0.5 * cap;
        const bool oversized = (double)grid.W * grid.H > (n_worlds > 1 ? 64.0 : 4096.0) && ((double)ex * 3 < grid.W && (double)ey * 3 < grid.H);
        if ((aliased || oversized) && !strip_on) bp_dirty = true;  // strip mode: a rebuild is a collective, keep the table
    }
    if (h_stats->nan_flag & 2u) return fail(BLOBS_ERR_NAN, "assertion failed: rotation is finite (physics.rs:471-474)");
    // strip mode: results are not valid after either of these. The statistics above are complete, and every rank sees the flag in
    // the same call (a rank that returned early would leave its neighbours waiting), so the call itself has run to its end.
    if (h_stats->nan_flag & 8u) return fail(BLOBS_ERR_CUDA, "strip exchange: a neighbour's message did not arrive in time (stale ghosts were used); results are invalid");
    if (h_stats->nan_flag & 4u) return fail(BLOBS_ERR_CAPACITY, "strip exchange: ghost / migration message or owned-body list overflowed (raise ghost_capacity / migrate_capacity); results are invalid");
    return BLOBS_OK;
}

int World::step(double delta, uint32_t n, BlobsStepStats* stats) {
    Span span("step");        // physics.rs:79
    if (collisions_enabled && use_spatial_hash && substeps > 0) return fail(BLOBS_ERR_SPATIAL_HASH, "spatial collisions not supported right now");
    int rc = flush();
    if (rc) return rc;
    if (topo_error) return fail(topo_error, topo_error_msg);
    if (nl_on) {
        if (collisions_enabled && !nl_collisions_were_on) nl_force_pending = true;   // the lists were not maintained meanwhile
        nl_collisions_were_on = collisions_enabled;
        if (nl_force_pending) {
            const unsigned int one = 1u;
            CU(cudaMemcpyAsync(&d_nlctl->force, &one, sizeof(one), cudaMemcpyHostToDevice, stream));
            nl_force_pending = false;
        }
    }
    DeviceStats init{};
    init.bb_min_x = init.bb_min_y = INT32_MAX;
    init.bb_max_x = init.bb_max_y = INT32_MIN;
    *h_stats = init;
    CU(cudaMemcpyAsync(d_stats, h_stats, sizeof(DeviceStats), cudaMemcpyHostToDevice, stream));
    if (strip_on) {
        rc = strip_rebuild_olist();
        if (rc) return rc;
        // arrivals are appended on the device: bound the launch for the whole call, rounded so the captured graph stays valid
        const size_t bound = (size_t)olaunch + 2 * (size_t)strip.mcap * substeps * n;
        olaunch_dim = (uint32_t)std::min<size_t>((bound + 32767) / 32768 * 32768, olist.cap);
    }
    CU(cudaEventRecord(ev_step0, stream));
    shadow_valid = false;
    for (uint32_t i = 0; i < n; ++i) {
        rc = run_step(substeps, (float)delta, i + 1 == n, !profiling || n == 1);  // physics.rs:80
        if (rc) return rc;
        time += delta;                           // physics.rs:81
    }
    return finish_stats(stats, n, n * substeps);
}

// Physics::fixed_step (physics.rs:84-99)
int World::fixed_step(double frame_time, BlobsStepStats* stats) {
    Span span("step");        // physics.rs:85
    if (collisions_enabled && use_spatial_hash && substeps > 0) return fail(BLOBS_ERR_SPATIAL_HASH, "spatial collisions not supported right now");
    int rc = flush();
    if (rc) return rc;
    if (topo_error) return fail(topo_error, topo_error_msg);
    if (nl_on) {
        if (collisions_enabled && !nl_collisions_were_on) nl_force_pending = true;   // the lists were not maintained meanwhile
        nl_collisions_were_on = collisions_enabled;
        if (nl_force_pending) {
            const unsigned int one = 1u;
            CU(cudaMemcpyAsync(&d_nlctl->force, &one, sizeof(one), cudaMemcpyHostToDevice, stream));
            nl_force_pending = false;
        }
    }
    DeviceStats init{};
    init.bb_min_x = init.bb_min_y = INT32_MAX;
    init.bb_max_x = init.bb_max_y = INT32_MIN;
    *h_stats = init;
    CU(cudaMemcpyAsync(d_stats, h_stats, sizeof(DeviceStats), cudaMemcpyHostToDevice, stream));
    if (strip_on) {
        rc = strip_rebuild_olist();
        if (rc) return rc;
        const size_t bound = (size_t)olaunch + 2 * (size_t)strip.mcap * substeps * 3;
        olaunch_dim = (uint32_t)std::min<size_t>((bound + 32767) / 32768 * 32768, olist.cap);
    }
    CU(cudaEventRecord(ev_step0, stream));
    shadow_valid = false;
    accumulator += frame_time;
    const double delta = 1.0 / 60.0;
    int max_steps = 3;
    uint32_t n = 0;
    while (accumulator >= delta && max_steps > 0) {
        rc = run_step(substeps, (float)delta, max_steps == 1 || !(accumulator - delta >= delta), !profiling);
        if (rc) return rc;
        accumulator -= delta;
        time += delta;
        max_steps -= 1;
        n++;
    }
    return finish_stats(stats, n, n * substeps);
}

// ---------------------------------------------------------------------------------------------- bulk IO
int World::download_bodies(BlobsBodyState* st, uint64_t* handles, size_t cap) {
    int rc = flush();
    if (rc) return rc;
    const size_t n = std::min(cap, bodies.slots());
    std::vector<float2> p(n), po(n), a(n), v(n), vr(n);
    std::vector<float> r(n), w(n), t(n);
    std::vector<uint8_t> hv(n);
    if (n && st) {
        CU(cudaMemcpyAsync(p.data(), pos.d, n * sizeof(float2), cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(po.data(), pos_old.d, n * sizeof(float2), cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(a.data(), acc.d, n * sizeof(float2), cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(v.data(), vel.d, n * sizeof(float2), cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(vr.data(), vreq.d, n * sizeof(float2), cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(r.data(), rot.d, n * sizeof(float), cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(w.data(), angvel.d, n * sizeof(float), cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(t.data(), torque.d, n * sizeof(float), cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(hv.data(), has_vreq.d, n, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
    }
    for (size_t s = 0; s < n; ++s) {
        const bool alive = bodies.alive[s];
        if (handles) handles[s] = alive ? bodies.handle_at((uint32_t)s) : 0;
        if (!st) continue;
        std::memset(&st[s], 0, sizeof(BlobsBodyState));
        if (!alive) continue;
        const HBody& b = hb[s];
        BlobsBodyState& o = st[s];
        o.position = {p[s].x, p[s].y}; o.position_old = {po[s].x, po[s].y}; o.acceleration = {a[s].x, a[s].y};
        o.calculated_velocity = {v[s].x, v[s].y}; o.velocity_request = {vr[s].x, vr[s].y}; o.has_velocity_request = hv[s];
        o.rotation = r[s]; o.angular_velocity = w[s]; o.torque = t[s];
        o.center_of_mass = b.com; o.scale = b.scale;
        o.calculated_mass = bmg.h[s].x; o.inertia = inertia.h[s]; o.gravity_mod = bmg.h[s].y;
        o.body_type = b.type; o.user_data_lo = b.ud_lo; o.user_data_hi = b.ud_hi;
    }
    return BLOBS_OK;
}

int World::download_colliders(BlobsColliderState* st, uint64_t* handles, size_t cap) {
    int rc = flush();
    if (rc) return rc;
    const size_t n = std::min(cap, cols.slots());
    std::vector<float2> a(n);
    std::vector<float> r(bodies.slots());
    if (n && st) {
        CU(cudaMemcpyAsync(a.data(), cabs.d, n * sizeof(float2), cudaMemcpyDeviceToHost, stream));
        if (!r.empty()) CU(cudaMemcpyAsync(r.data(), rot.d, r.size() * sizeof(float), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
    }
    for (size_t s = 0; s < n; ++s) {
        const bool alive = cols.alive[s];
        if (handles) handles[s] = alive ? cols.handle_at((uint32_t)s) : 0;
        if (!st) continue;
        std::memset(&st[s], 0, sizeof(BlobsColliderState));
        if (!alive) continue;
        st[s].desc = hc[s].desc;
        st[s].desc.absolute_transform = live_snapshot((uint32_t)s, a[s], r);
        st[s].parent = hc[s].parent;
    }
    return BLOBS_OK;
}

// collider.absolute_transform as the reference holds it: the caller's value until a substep has run, then
// rbd.transform() * offset (physics.rs:360-366): the translation is what the device keeps (bit-exact), the matrix part is
// only ever read by debug output, so it is recomposed here: M(rot) * offset.matrix2 (glam column convention).
BlobsAffine2 World::live_snapshot(uint32_t s, float2 t, const std::vector<float>& rot_host) const {
    BlobsAffine2 out = hc[s].desc.absolute_transform;
    out.translation = {t.x, t.y};
    if (hc[s].born_epoch < snap_epoch && bodies.valid(hc[s].parent)) {
        const uint32_t b = h_slot(hc[s].parent);
        const float th = b < rot_host.size() ? rot_host[b] : 0.f;
        const float sn = std::sin(th), cs = std::cos(th);
        const BlobsAffine2& o = hc[s].desc.offset;
        // M * v = x_axis * v.x + y_axis * v.y with x_axis = (cos, sin), y_axis = (-sin, cos)
        out.x_axis = {cs * o.x_axis.x + (-sn) * o.x_axis.y, sn * o.x_axis.x + cs * o.x_axis.y};
        out.y_axis = {cs * o.y_axis.x + (-sn) * o.y_axis.y, sn * o.y_axis.x + cs * o.y_axis.y};
    }
    return out;
}

// Physics::debug_data (physics.rs:479-481, debug.rs:34-91): arena iteration order = ascending slot, free slots skipped.
int World::debug_counts(BlobsDebugCounts* out) const {
    out->bodies = bodies.len;
    out->joints = joints.len;
    out->colliders = cols.len;
    out->springs = springs.len;
    return BLOBS_OK;
}

int World::debug_data(float* body_xform, float* joint_ab, float* col_xform, float* col_radius, float* spring_ab, const BlobsDebugCounts* caps) {
    int rc = flush();
    if (rc) return rc;
    const size_t nb = bodies.slots(), nc = cols.slots();
    std::vector<float2> p(nb), a(nc);
    std::vector<float> r(nb);
    if (nb) {
        CU(cudaMemcpyAsync(p.data(), pos.d, nb * sizeof(float2), cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(r.data(), rot.d, nb * sizeof(float), cudaMemcpyDeviceToHost, stream));
    }
    if (nc) CU(cudaMemcpyAsync(a.data(), cabs.d, nc * sizeof(float2), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    const float nanv = std::nanf("");
    if (body_xform) {   // DebugRigidBody { transform: Affine2::from_angle_translation(rotation, position) }
        size_t k = 0;
        for (size_t s = 0; s < nb && k < caps->bodies; ++s) {
            if (!bodies.alive[s]) continue;
            const float sn = std::sin(r[s]), cs = std::cos(r[s]);
            float* o = body_xform + 6 * k++;
            o[0] = cs; o[1] = sn; o[2] = -sn; o[3] = cs; o[4] = p[s].x; o[5] = p[s].y;
        }
    }
    auto endpoints = [&](float* out, size_t cap, const HostArena& arena, auto&& ab_of) {
        size_t k = 0;
        for (size_t s = 0; s < arena.slots() && k < cap; ++s) {
            if (!arena.alive[s]) continue;
            const auto ab = ab_of(s);
            float* o = out + 4 * k++;
            // the reference indexes the arena directly and panics on a removed body; here the endpoint reads as NaN
            const bool va = bodies.valid(ab.first), vb = bodies.valid(ab.second);
            o[0] = va ? p[h_slot(ab.first)].x : nanv; o[1] = va ? p[h_slot(ab.first)].y : nanv;
            o[2] = vb ? p[h_slot(ab.second)].x : nanv; o[3] = vb ? p[h_slot(ab.second)].y : nanv;
        }
    };
    if (joint_ab) endpoints(joint_ab, caps->joints, joints, [&](size_t s) { return std::make_pair(hj[s].a, hj[s].b); });
    if (spring_ab) endpoints(spring_ab, caps->springs, springs, [&](size_t s) { return std::make_pair(hs[s].a, hs[s].b); });
    if (col_xform || col_radius) {   // DebugCollider { transform: collider.absolute_transform, radius: ball.radius }
        size_t k = 0;
        for (size_t s = 0; s < nc && k < caps->colliders; ++s) {
            if (!cols.alive[s]) continue;
            if (col_xform) {
                const BlobsAffine2 t = live_snapshot((uint32_t)s, a[s], r);
                float* o = col_xform + 6 * k;
                o[0] = t.x_axis.x; o[1] = t.x_axis.y; o[2] = t.y_axis.x; o[3] = t.y_axis.y; o[4] = t.translation.x; o[5] = t.translation.y;
            }
            if (col_radius) col_radius[k] = hc[s].desc.shape_radius;
            ++k;
        }
    }
    return BLOBS_OK;
}

int World::read_body_vec(int which, float* xy, size_t cap) {
    int rc = flush();
    if (rc) return rc;
    const size_t n = std::min(cap, bodies.slots());
    if (!n) return BLOBS_OK;
    const float2* src = which == 0 ? pos.d : vel.d;
    CU(cudaMemcpyAsync(xy, src, n * sizeof(float2), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    return BLOBS_OK;
}

int World::apply_forces(const float* f, size_t cap) {
    int rc = flush();
    if (rc) return rc;
    const size_t n = std::min(cap, bodies.slots());
    if (!n) return BLOBS_OK;
    CU(d_forces.ensure(n, stream));
    CU(cudaMemcpyAsync(d_forces.d, f, n * sizeof(float2), cudaMemcpyHostToDevice, stream));
    BLOBS_LAUNCH(cdiv(n, 256), 256, 0, stream, k_apply_forces)(body_arrays(), d_forces.d, (uint32_t)n);
    launches++;
    CU(cudaGetLastError());
    return BLOBS_OK;
}

// ---- pipelined host I/O -------------------------------------------------------------------------------------------------
// blobs_apply_forces / blobs_read_body_positions put their PCIe copy on the compute stream, so a frame costs
// H2D + step + D2H back to back. The calls below move the copies to one stream per direction: the forces of step i+1 travel
// while step i computes, and the positions of step i (snapshotted device-to-device first, so step i+1 may overwrite them)
// travel while step i+1 computes. Ordering is by events only; results are identical to the synchronous calls.
int World::io_init() {
    if (io_ready) return BLOBS_OK;
    CU(cudaStreamCreateWithFlags(&s_h2d, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CU(cudaEventCreateWithFlags(&ev_up_done[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ev_up_free[i], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&ev_snap_ready, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ev_snap_free, cudaEventDisableTiming));
    io_ready = true;
    return BLOBS_OK;
}

// starts the host->device copy of one batch of per-slot forces (double-buffered on the device); returns at once
int World::forces_upload_async(const float* f, size_t cap) {
    int rc = io_init();
    if (rc) return rc;
    if (up_pending >= 0) return fail(BLOBS_ERR_INVALID, "blobs_forces_upload_async: the previous batch has not been applied yet");
    const size_t n = std::min(cap, bodies.slots());
    const int k = up_next;
    up_next ^= 1;
    if (n) {
        if (up_used[k]) CU(cudaStreamWaitEvent(s_h2d, ev_up_free[k], 0));   // the kernel that read this buffer two batches ago
        CU(d_forces_up[k].ensure(n, s_h2d));
        CU(cudaMemcpyAsync(d_forces_up[k].d, f, n * sizeof(float2), cudaMemcpyHostToDevice, s_h2d));
    }
    CU(cudaEventRecord(ev_up_done[k], s_h2d));
    up_pending = k;
    up_n = n;
    return BLOBS_OK;
}

// RigidBody::apply_force (rigid_body.rs:155-160) for every slot of the uploaded batch; the compute stream waits for the copy
int World::apply_forces_uploaded() {
    if (up_pending < 0 || up_indexed) return fail(BLOBS_ERR_INVALID, "blobs_apply_forces_uploaded: no uploaded batch");
    int rc = flush();
    if (rc) return rc;
    const int k = up_pending;
    up_pending = -1;
    const size_t n = std::min(up_n, bodies.slots());
    CU(cudaStreamWaitEvent(stream, ev_up_done[k], 0));
    if (n) {
        BLOBS_LAUNCH(cdiv(n, 256), 256, 0, stream, k_apply_forces)(body_arrays(), d_forces_up[k].d, (uint32_t)n);
        launches++;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(ev_up_free[k], stream));
    up_used[k] = true;
    return BLOBS_OK;
}

// positions of all body slots -> host, without blocking: device-to-device snapshot on the compute stream, PCIe copy on its own
// stream. `xy` must stay untouched until blobs_io_sync returns.
int World::read_positions_async(float* xy, size_t cap) {
    int rc = io_init();
    if (rc) return rc;
    rc = flush();
    if (rc) return rc;
    const size_t n = std::min(cap, bodies.slots());
    if (!n) return BLOBS_OK;
    if (snap_used) CU(cudaStreamWaitEvent(stream, ev_snap_free, 0));   // the previous snapshot is still on its way to the host
    CU(d_pos_snap.ensure(n, stream));
    CU(cudaMemcpyAsync(d_pos_snap.d, pos.d, n * sizeof(float2), cudaMemcpyDeviceToDevice, stream));
    CU(cudaEventRecord(ev_snap_ready, stream));
    CU(cudaStreamWaitEvent(s_d2h, ev_snap_ready, 0));
    CU(cudaMemcpyAsync(xy, d_pos_snap.d, n * sizeof(float2), cudaMemcpyDeviceToHost, s_d2h));
    CU(cudaEventRecord(ev_snap_free, s_d2h));
    snap_used = true;
    return BLOBS_OK;
}

// ---- the same pipeline for the distributed (strip) host I/O: indexed forces in, compact (slot, position) list of the owned bodies out
int World::forces_indexed_upload_async(const uint32_t* slots, const float* f, size_t n) {
    int rc = io_init();
    if (rc) return rc;
    if (up_pending >= 0) return fail(BLOBS_ERR_INVALID, "blobs_forces_indexed_upload_async: the previous batch has not been applied yet");
    const int k = up_next;
    up_next ^= 1;
    if (n) {
        if (up_used[k]) CU(cudaStreamWaitEvent(s_h2d, ev_up_free[k], 0));   // the kernel that read this buffer two batches ago
        CU(d_fslots_up[k].ensure(n, s_h2d));
        CU(d_forces_up[k].ensure(n, s_h2d));
        CU(cudaMemcpyAsync(d_fslots_up[k].d, slots, n * sizeof(uint32_t), cudaMemcpyHostToDevice, s_h2d));
        CU(cudaMemcpyAsync(d_forces_up[k].d, f, n * sizeof(float2), cudaMemcpyHostToDevice, s_h2d));
    }
    CU(cudaEventRecord(ev_up_done[k], s_h2d));
    up_pending = k;
    up_n = n;
    up_indexed = true;
    return BLOBS_OK;
}

int World::apply_forces_indexed_uploaded() {
    if (up_pending < 0 || !up_indexed) return fail(BLOBS_ERR_INVALID, "blobs_apply_forces_indexed_uploaded: no uploaded indexed batch");
    int rc = flush();
    if (rc) return rc;
    const int k = up_pending;
    up_pending = -1;
    up_indexed = false;
    CU(cudaStreamWaitEvent(stream, ev_up_done[k], 0));
    if (up_n) {
        BLOBS_LAUNCH(cdiv(up_n, 256), 256, 0, stream, k_apply_forces_indexed)(body_arrays(), strip_on ? d_owned.d : nullptr, d_fslots_up[k].d, d_forces_up[k].d, (uint32_t)up_n,
                                                                            (uint32_t)bodies.slots());
        launches++;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(ev_up_free[k], stream));
    up_used[k] = true;
    return BLOBS_OK;
}

// compaction on the compute stream into a snapshot pair, PCIe copies (count, slots, positions) on the device-to-host stream
int World::read_owned_positions_async(uint32_t* slots, float* xy, uint32_t* n_out, size_t cap) {
    int rc = io_init();
    if (rc) return rc;
    rc = flush();
    if (rc) return rc;
    const size_t nb = bodies.slots();
    if (!d_ocount_snap) CU(cudaMalloc(&d_ocount_snap, sizeof(unsigned int)));
    // entries that can be valid: the owned-list bound of the step just run (strips), every body slot otherwise
    const size_t m = std::min<size_t>(cap, strip_on ? std::max<size_t>(olaunch_dim, 1) : std::max<size_t>(nb, 1));
    if (snap_used) CU(cudaStreamWaitEvent(stream, ev_snap_free, 0));   // the previous snapshot is still on its way to the host
    CU(d_oslots_snap.ensure(std::max<size_t>(m, 1), stream));
    CU(d_oxy_snap.ensure(std::max<size_t>(m, 1), stream));
    CU(cudaMemsetAsync(d_ocount_snap, 0, sizeof(unsigned int), stream));
    if (nb && m) {
        BLOBS_LAUNCH(cdiv(nb, 256), 256, 0, stream, k_compact_owned)(body_arrays(), strip_on ? d_owned.d : nullptr, (uint32_t)nb, (uint32_t)std::min<size_t>(m, 0xffffffffu),
                                                           d_ocount_snap, d_oslots_snap.d, d_oxy_snap.d);
        launches++;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(ev_snap_ready, stream));
    CU(cudaStreamWaitEvent(s_d2h, ev_snap_ready, 0));
    CU(cudaMemcpyAsync(n_out, d_ocount_snap, sizeof(unsigned int), cudaMemcpyDeviceToHost, s_d2h));
    if (m) {
        CU(cudaMemcpyAsync(slots, d_oslots_snap.d, m * sizeof(uint32_t), cudaMemcpyDeviceToHost, s_d2h));
        CU(cudaMemcpyAsync(xy, d_oxy_snap.d, m * sizeof(float2), cudaMemcpyDeviceToHost, s_d2h));
    }
    CU(cudaEventRecord(ev_snap_free, s_d2h));
    snap_used = true;
    return BLOBS_OK;
}

int World::io_sync() {
    if (!io_ready) return BLOBS_OK;
    CU(cudaStreamSynchronize(s_h2d));
    CU(cudaStreamSynchronize(s_d2h));
    return BLOBS_OK;
}

int World::download_cell_coords(int32_t* cx, int32_t* cy, size_t cap) {
    int rc = flush();
    if (rc) return rc;
    const size_t n = std::min(cap, cols.slots());
    if (!n) return BLOBS_OK;
    CU(d_cellx.ensure(n, stream));
    CU(d_celly.ensure(n, stream));
    BLOBS_LAUNCH(cdiv(n, 256), 256, 0, stream, k_cell_coords)(cabs.d, cell_size, (uint32_t)n, d_cellx.d, d_celly.d);
    launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(cx, d_cellx.d, n * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CU(cudaMemcpyAsync(cy, d_celly.d, n * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    for (size_t s = 0; s < n; ++s)
        if (!cols.alive[s]) cx[s] = cy[s] = 0;
    return BLOBS_OK;
}

// ---------------------------------------------------------------------------------------------- scene queries
int World::query_circles(size_t n, const float* centre_xy, const float* radius, const BlobsQueryFilter* filter, uint64_t* offsets, uint64_t* hits,
                         size_t hit_cap, size_t* n_hits) {
    if (strip_on) return fail(BLOBS_ERR_INVALID, "scene queries are not available on a strip-decomposed world (each rank holds one strip)");
    if (n > 0xffffffffull) return fail(BLOBS_ERR_INVALID, "too many queries in one call");
    int rc = flush();   // also (re)builds the table if colliders were inserted or moved since the last step
    if (rc) return rc;
    if (n_hits) *n_hits = 0;
    for (size_t q = 0; q <= n; ++q) offsets[q] = 0;
    if (!n || !cols.slots() || !n_active_cols) return BLOBS_OK;
    QueryFilterDev F{};
    F.exclude_col = F.exclude_body = NO_SLOT;
    if (filter) {
        F.flags = filter->flags;
        F.has_groups = filter->has_groups ? 1u : 0u;
        F.memb = filter->memberships;
        F.filt = filter->filter;
        if (filter->exclude_collider && cols.valid(filter->exclude_collider)) F.exclude_col = h_slot(filter->exclude_collider);
        if (filter->exclude_rigid_body && bodies.valid(filter->exclude_rigid_body)) F.exclude_body = h_slot(filter->exclude_rigid_body);
        if (filter->batch_world >= n_worlds) return fail(BLOBS_ERR_INVALID, "query: batch world id out of range");
        F.wbase = filter->batch_world * grid.ncells;
    }
    Broadphase bp{};
    bp.hot = cur_is_a ? hot_a.d : hot_b.d;
    bp.tab = cur_is_a ? tab_a.d : tab_b.d;
    bp.tab_next = nullptr;
    bp.tile_next = nullptr;
    if (nl_on) {   // list pipeline: the grid dates from the last list rebuild - sort the current snapshots into it first
        rc = nl_rebuild_now();
        if (rc) return rc;
        bp.hot = hot_a.d;
        bp.tab = (h_nlctl->parity & 1u) ? tab_b.d : tab_a.d;
    }
    CU(d_qcentre.ensure(n, stream)); CU(d_qradius.ensure(n, stream)); CU(d_qcount.ensure(n, stream)); CU(d_qoff.ensure(n, stream));
    CU(cudaMemcpyAsync(d_qcentre.d, centre_xy, n * sizeof(float2), cudaMemcpyHostToDevice, stream));
    CU(cudaMemcpyAsync(d_qradius.d, radius, n * sizeof(float), cudaMemcpyHostToDevice, stream));
    const BodyArrays B = body_arrays();
    const ColliderArrays C = col_arrays();
    BLOBS_LAUNCH(cdiv(n, 128), 128, 0, stream, k_query)(grid, bp, C, B, d_qcentre.d, d_qradius.d, (uint32_t)n, F, nullptr, d_qcount.d, nullptr);
    launches++;
    CU(cudaGetLastError());
    std::vector<uint32_t> cnt(n), off(n);
    CU(cudaMemcpyAsync(cnt.data(), d_qcount.d, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    uint64_t total = 0;
    for (size_t q = 0; q < n; ++q) { offsets[q] = total; off[q] = (uint32_t)total; total += cnt[q]; }
    offsets[n] = total;
    if (n_hits) *n_hits = (size_t)total;
    if (total > 0xffffffffull) return fail(BLOBS_ERR_CAPACITY, "query: more than 2^32 hits in one call");
    if (total > hit_cap) return fail(BLOBS_ERR_CAPACITY, "query: hit buffer too small (n_hits holds the required capacity)");
    if (!total) return BLOBS_OK;
    CU(d_qhits.ensure((size_t)total, stream));
    CU(cudaMemcpyAsync(d_qoff.d, off.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    BLOBS_LAUNCH(cdiv(n, 128), 128, 0, stream, k_query)(grid, bp, C, B, d_qcentre.d, d_qradius.d, (uint32_t)n, F, d_qoff.d, d_qcount.d, d_qhits.d);
    launches++;
    CU(cudaGetLastError());
    std::vector<uint32_t> slots((size_t)total);
    CU(cudaMemcpyAsync(slots.data(), d_qhits.d, (size_t)total * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    for (size_t q = 0; q < n; ++q) {
        std::sort(slots.begin() + (ptrdiff_t)offsets[q], slots.begin() + (ptrdiff_t)offsets[q + 1]);
        for (uint64_t i = offsets[q]; i < offsets[q + 1]; ++i) hits[i] = cols.handle_at(slots[i]);
    }
    return BLOBS_OK;
}

// ---------------------------------------------------------------------------------------------- recording
int World::record_contacts(int mode, size_t cap) {
    if (mode < 0 || mode > 2) return fail(BLOBS_ERR_INVALID, "bad record mode");
    if (strip_on && mode == BLOBS_RECORD_EVENTS)   // (the other order is refused by strip_configure)
        return fail(BLOBS_ERR_INVALID, "event recording is not supported on a strip-decomposed world (DESIGN.md 8.1)");
    CU(cudaStreamSynchronize(stream));
    if ((mode == BLOBS_RECORD_EVENTS) != (rec_mode == BLOBS_RECORD_EVENTS)) topo_dirty = bp_dirty = true;  // CF_COLD depends on it
    rec_mode = mode;
    rec_cap = mode ? std::max<size_t>(cap, 1) : 0;
    if (mode) {
        CU(rec_pairs.ensure(rec_cap, stream));
        if (mode == BLOBS_RECORD_EVENTS) CU(rec_vels.ensure(rec_cap, stream));
        CU(d_sub_end.ensure(8192, stream));
    }
    CU(cudaMemsetAsync(d_rec_count, 0, sizeof(unsigned long long), stream));
    sub_recorded = 0;
    rec_drained = 0;
    return BLOBS_OK;
}

int World::pairs_drain(uint32_t* a, uint32_t* b, size_t cap, size_t* n, uint64_t* sub_end, size_t sub_cap, size_t* n_sub) {
    unsigned long long cnt = 0;
    CU(cudaMemcpyAsync(&cnt, d_rec_count, sizeof(cnt), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    const size_t have = (size_t)std::min<unsigned long long>(cnt, rec_cap);
    const size_t m = std::min(have, cap);
    std::vector<uint2> tmp(m);
    if (m) {
        CU(cudaMemcpyAsync(tmp.data(), rec_pairs.d, m * sizeof(uint2), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
    }
    for (size_t i = 0; i < m; ++i) {
        if (a) a[i] = tmp[i].x;
        if (b) b[i] = tmp[i].y;
    }
    if (n) *n = have;
    const size_t ns = std::min<size_t>(sub_recorded, sub_cap);
    if (sub_end && ns) {
        std::vector<unsigned long long> se(ns);
        CU(cudaMemcpyAsync(se.data(), d_sub_end.d, ns * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        for (size_t i = 0; i < ns; ++i) sub_end[i] = se[i];
    }
    if (n_sub) *n_sub = sub_recorded;
    CU(cudaMemsetAsync(d_rec_count, 0, sizeof(unsigned long long), stream));
    sub_recorded = 0;
    return BLOBS_OK;
}

// A drain may be partial (cap smaller than what was recorded): *n reports how many events were still pending BEFORE this call,
// the first min(*n, cap) of them are returned and consumed, the rest stay for the next call (the channel of the reference,
// physics.rs:22-23, never loses an event either).
int World::events_drain(BlobsCollisionEvent* buf, size_t cap, size_t* n) {
    if (rec_mode != BLOBS_RECORD_EVENTS) return fail(BLOBS_ERR_INVALID, "event recording is not enabled");
    unsigned long long cnt = 0;
    CU(cudaMemcpyAsync(&cnt, d_rec_count, sizeof(cnt), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    const size_t recorded = (size_t)std::min<unsigned long long>(cnt, rec_cap);
    if (rec_drained > recorded) rec_drained = recorded;
    const size_t have = recorded - rec_drained;
    const size_t m = std::min(have, cap);
    std::vector<uint2> pr(m);
    std::vector<float4> vl(m);
    if (m) {
        CU(cudaMemcpyAsync(pr.data(), rec_pairs.d + rec_drained, m * sizeof(uint2), cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(vl.data(), rec_vels.d + rec_drained, m * sizeof(float4), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
    }
    for (size_t i = 0; i < m; ++i) {
        buf[i].col_handle_a = cols.handle_at(pr[i].x);
        buf[i].col_handle_b = cols.handle_at(pr[i].y);
        buf[i].impact_vel_a = {vl[i].x, vl[i].y};
        buf[i].impact_vel_b = {vl[i].z, vl[i].w};
    }
    if (n) *n = have;
    rec_drained += m;
    if (rec_drained == recorded) {   // everything consumed: recording restarts at the beginning of the buffer
        CU(cudaMemsetAsync(d_rec_count, 0, sizeof(unsigned long long), stream));
        rec_drained = 0;
        sub_recorded = 0;
    }
    return BLOBS_OK;
}

// ---------------------------------------------------------------------------------------------- strip decomposition
namespace {
struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool load(std::string* err) {
        if (lib) return true;
        // resolved at run time so that single-GPU use has no NCCL dependency; inside a torch process this finds torch's copy
#ifdef BLOBS_EMU   // host-compiled test build (tests/emu): a socket-based stand-in named by the test, never the real library
        if (const char* fake = std::getenv("BLOBS_EMU_NCCL_LIB")) lib = dlopen(fake, RTLD_NOW | RTLD_GLOBAL);
#else
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
#endif
        if (!lib) { if (err) *err = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
#define BLOBS_NCCL_SYM(f) f = reinterpret_cast<decltype(f)>(dlsym(lib, "nccl" #f)); if (!f) { if (err) *err = "libnccl lacks nccl" #f; return false; }
        BLOBS_NCCL_SYM(GetUniqueId) BLOBS_NCCL_SYM(CommInitRank) BLOBS_NCCL_SYM(CommDestroy) BLOBS_NCCL_SYM(Send) BLOBS_NCCL_SYM(Recv)
        BLOBS_NCCL_SYM(GroupStart) BLOBS_NCCL_SYM(GroupEnd) BLOBS_NCCL_SYM(GetErrorString)
#undef BLOBS_NCCL_SYM
        return true;
    }
};
NcclApi g_nccl;
}  // namespace

#define NC(call)                                                                                      \
    do {                                                                                              \
        ncclResult_t r__ = (call);                                                                    \
        if (r__ != ncclSuccess) return fail(BLOBS_ERR_CUDA, std::string("NCCL error: ") + g_nccl.GetErrorString(r__) + " in " #call); \
    } while (0)

int World::strip_unique_id(uint8_t* out128, std::string* err) {
    if (!g_nccl.load(err)) return BLOBS_ERR_CUDA;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) { if (err) *err = "ncclGetUniqueId failed"; return BLOBS_ERR_CUDA; }
    std::memcpy(out128, &id, 128);
    return BLOBS_OK;
}

int World::strip_configure(int rank, int nranks, float x_lo, float x_hi, const uint8_t* id128, uint32_t gcap, uint32_t mcap) {
    if (nranks < 1 || rank < 0 || rank >= nranks || !(x_lo < x_hi)) return fail(BLOBS_ERR_INVALID, "bad strip arguments");
    int rc = flush();
    if (rc) return rc;
    if (topo_error) return fail(topo_error, topo_error_msg);
    if (n_multi || n_springs_live || n_joints_live || n_worlds > 1 || rec_mode == BLOBS_RECORD_EVENTS)
        return fail(BLOBS_ERR_INVALID, "strip decomposition supports single-collider bodies without springs/joints/events (DESIGN.md 8.1)");
    if (nranks > 1) {
        std::string e;
        if (!g_nccl.load(&e)) return fail(BLOBS_ERR_CUDA, e);
        ncclUniqueId id;
        std::memcpy(&id, id128, 128);
        ncclComm_t comm = nullptr;
        NC(g_nccl.CommInitRank(&comm, nranks, id, rank));
        nccl_comm = comm;
        g_nccl_destroy = [](void* c) { g_nccl.CommDestroy(static_cast<ncclComm_t>(c)); };
    }
    s_rank = rank;
    s_nranks = nranks;
    strip.x_lo = x_lo;
    strip.x_hi = x_hi;
    strip.has_left = rank > 0;
    strip.has_right = rank + 1 < nranks;
    strip.rmax = r_max;
    strip.gcap = std::max<uint32_t>(gcap, 256);
    strip.mcap = std::max<uint32_t>(mcap, 64);
    msg_bytes = strip_msg_bytes(strip.gcap, strip.mcap);
    for (int i = 0; i < 4; ++i) {
        CU(cudaMalloc(&msg[i], msg_bytes));
        CU(cudaMemsetAsync(msg[i], 0, msg_bytes, stream));
    }
    cur_recv[0] = msg[2];
    cur_recv[1] = msg[3];
    p2p_on = false;
    if (p2p_request) {
        rc = strip_p2p_setup();   // a one-rank strip world has nobody to exchange with: the same code paths with no peer (diagnostics)
        if (rc) return rc;
        if (p2p_on && list_mode != 0) {   // neighbour lists on strips need the peer mappings as well
            rc = nls_setup();
            if (rc) return rc;
        }
    }
    CU(gcell.ensure(2 * (size_t)strip.gcap, stream));
    const size_t nb = bodies.slots(), nc = cols.slots();
    CU(d_owned.ensure(std::max<size_t>(nb, 1), stream));
    CU(d_cowned.ensure(std::max<size_t>(nc, 1), stream));
    CU(cudaMemsetAsync(d_cowned.d, 0, d_cowned.cap, stream));
    if (nb) {
        BLOBS_LAUNCH(cdiv(nb, 256), 256, 0, stream, k_strip_init_owned)(body_arrays(), col_arrays(), strip, d_owned.d, d_cowned.d, (uint32_t)nb);
        launches++;
        CU(cudaGetLastError());
    }
    strip_on = true;
    CU(olist.ensure(std::max<size_t>(nb, 1), stream));
    CU(opos.ensure(std::max<size_t>(nb, 1), stream));
    if (!d_ocount) CU(cudaMalloc(&d_ocount, sizeof(uint32_t)));
    rc = strip_rebuild_olist();
    if (rc) return rc;
    bp_dirty = true;
    return rebuild_broadphase();
}

StripView World::strip_view() {
    StripView v{};
    if (strip_on) {
        v.olist = olist.d;
        v.ocount = d_ocount;
        v.send_l = msg[0];
        v.send_r = msg[1];
        v.S = strip;
        v.cowned = d_cowned.d;
    }
    return v;
}

// compact list of owned bodies in (warp-granular) slot order; once per blobs_step* call, so released entries never pile up
int World::strip_rebuild_olist() {
    const size_t nb = bodies.slots();
    CU(cudaMemsetAsync(d_ocount, 0, sizeof(uint32_t), stream));
    if (nb) {
        BLOBS_LAUNCH(cdiv(nb, 256), 256, 0, stream, k_strip_build_olist)(d_owned.d, (uint32_t)nb, olist.d, d_ocount, opos.d);
        launches++;
        CU(cudaGetLastError());
    }
    uint32_t cnt = 0;
    CU(cudaMemcpyAsync(&cnt, d_ocount, sizeof(cnt), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    olaunch = cnt;
    olaunch_dim = olaunch;
    return BLOBS_OK;
}

// Peer-memory exchange set-up: allocate my receive block, swap IPC handles with both neighbours (one grouped ncclSend/ncclRecv
// of 64 bytes - NCCL is only the rendezvous here), map their blocks, then agree across ALL ranks (nranks - 1 rounds of
// neighbour min) whether everybody succeeded: the exchange is either peer stores on every rank or NCCL on every rank.
int World::strip_p2p_setup() {
    ncclComm_t comm = static_cast<ncclComm_t>(nccl_comm);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    p2p_stride = (msg_bytes + 255) / 256 * 256;
    int ok = 1;
    if (blobs_ipc_alloc(&p2p_block, 4 * p2p_stride) != cudaSuccess) { cudaGetLastError(); p2p_block = nullptr; ok = 0; }
    cudaIpcMemHandle_t hmine{}, hpeer[2]{};
    if (ok) {
        CU(cudaMemsetAsync(p2p_block, 0, 4 * p2p_stride, stream));
        if (cudaIpcGetMemHandle(&hmine, p2p_block) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    }
    if (!d_push_done) { CU(cudaMalloc(&d_push_done, 2 * sizeof(unsigned int))); CU(cudaMemsetAsync(d_push_done, 0, 2 * sizeof(unsigned int), stream)); }
    if (s_nranks <= 1) {
        p2p_on = ok != 0;
        return BLOBS_OK;
    }
    char* d_h = nullptr;   // [mine | from left | from right] handles, then [mine | left | right] ok words
    CU(cudaMalloc(&d_h, 3 * 64 + 3 * sizeof(int)));
    int* d_ok = reinterpret_cast<int*>(d_h + 3 * 64);
    CU(cudaMemcpyAsync(d_h, &hmine, 64, cudaMemcpyHostToDevice, stream));
    auto swap = [&](const void* send, void* from_l, void* from_r, size_t n) -> int {
        NC(g_nccl.GroupStart());
        if (strip.has_right) { NC(g_nccl.Send(send, n, ncclInt8, s_rank + 1, comm, stream)); NC(g_nccl.Recv(from_r, n, ncclInt8, s_rank + 1, comm, stream)); }
        if (strip.has_left) { NC(g_nccl.Send(send, n, ncclInt8, s_rank - 1, comm, stream)); NC(g_nccl.Recv(from_l, n, ncclInt8, s_rank - 1, comm, stream)); }
        NC(g_nccl.GroupEnd());
        CU(cudaStreamSynchronize(stream));
        return BLOBS_OK;
    };
    int rc = swap(d_h, d_h + 64, d_h + 128, 64);
    if (rc) return rc;
    CU(cudaMemcpy(hpeer, d_h + 64, 128, cudaMemcpyDeviceToHost));
    for (int side = 0; side < 2 && ok; ++side) {
        if (!(side ? strip.has_right : strip.has_left)) continue;
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, hpeer[side], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
        p2p_peer[side] = static_cast<char*>(q);
    }
    for (int round = 0; round + 1 < s_nranks; ++round) {   // global AND of `ok` along the chain of strips
        int h3[3] = {ok, 1, 1};
        CU(cudaMemcpy(d_ok, h3, sizeof(h3), cudaMemcpyHostToDevice));
        rc = swap(d_ok, d_ok + 1, d_ok + 2, sizeof(int));
        if (rc) return rc;
        CU(cudaMemcpy(h3, d_ok, sizeof(h3), cudaMemcpyDeviceToHost));
        ok = std::min(h3[0], std::min(h3[1], h3[2]));
    }
    cudaFree(d_h);
    p2p_on = ok != 0;
    return BLOBS_OK;
}

// List pipeline on strips: the two slot-indexed snapshot arrays and a small flag block become CUDA-IPC memory; every rank maps
// every other rank's flag block (the rebuild decision combines all ranks' numbers) and its neighbours' snapshot arrays (ghost
// records are stored there directly). All-or-nothing across ranks, like the peer-memory exchange itself.
int World::nls_setup() {
    nls_ready = false;
    if (s_nranks > NL_MAX_RANKS || !p2p_on) return BLOBS_OK;
    ncclComm_t comm = static_cast<ncclComm_t>(nccl_comm);
    struct Card { cudaIpcMemHandle_t snap, flags; int ok; int pad[15]; };
    static_assert(sizeof(Card) == 192, "Card layout");
    Card mine{};
    mine.ok = 1;
    nls_snap_cap = std::max<size_t>(cabs.cap, 1);
    const size_t fbytes = 2 * (size_t)NL_MAX_RANKS * sizeof(NlFlag);
    char* fl = nullptr;
    if (blobs_ipc_alloc(&nls_snap_block, 2 * nls_snap_cap * sizeof(float4)) != cudaSuccess) { cudaGetLastError(); nls_snap_block = nullptr; mine.ok = 0; }
    if (mine.ok && blobs_ipc_alloc(&fl, fbytes) != cudaSuccess) { cudaGetLastError(); fl = nullptr; mine.ok = 0; }
    nls_flags = reinterpret_cast<NlFlag*>(fl);
    if (mine.ok) {
        CU(cudaMemsetAsync(nls_snap_block, 0, 2 * nls_snap_cap * sizeof(float4), stream));
        CU(cudaMemsetAsync(fl, 0, fbytes, stream));
        if (cudaIpcGetMemHandle(&mine.snap, nls_snap_block) != cudaSuccess || cudaIpcGetMemHandle(&mine.flags, fl) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
    }
    char* d_cards = nullptr;
    CU(cudaMalloc(&d_cards, (size_t)s_nranks * sizeof(Card)));
    std::vector<Card> cards(s_nranks);
    auto allgather = [&]() -> int {   // grouped point-to-point: NCCL is only the rendezvous here
        CU(cudaMemcpyAsync(d_cards + (size_t)s_rank * sizeof(Card), &mine, sizeof(Card), cudaMemcpyHostToDevice, stream));
        CU(cudaStreamSynchronize(stream));
        if (s_nranks > 1) {
            NC(g_nccl.GroupStart());
            for (int r = 0; r < s_nranks; ++r) {
                if (r == s_rank) continue;
                NC(g_nccl.Send(d_cards + (size_t)s_rank * sizeof(Card), sizeof(Card), ncclInt8, r, comm, stream));
                NC(g_nccl.Recv(d_cards + (size_t)r * sizeof(Card), sizeof(Card), ncclInt8, r, comm, stream));
            }
            NC(g_nccl.GroupEnd());
        }
        CU(cudaStreamSynchronize(stream));
        CU(cudaMemcpy(cards.data(), d_cards, (size_t)s_nranks * sizeof(Card), cudaMemcpyDeviceToHost));
        return BLOBS_OK;
    };
    int rc = allgather();
    if (rc) return rc;
    int all_ok = 1;
    for (const Card& c : cards) all_ok = std::min(all_ok, c.ok);
    if (all_ok) {
        for (int r = 0; r < s_nranks && mine.ok; ++r) {
            if (r == s_rank) { nls_peer_flags[r] = nls_flags; continue; }
            void* q = nullptr;
            if (cudaIpcOpenMemHandle(&q, cards[r].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; break; }
            nls_peer_flags[r] = static_cast<NlFlag*>(q);
        }
        for (int side = 0; side < 2 && mine.ok; ++side) {
            if (!(side ? strip.has_right : strip.has_left)) continue;
            void* q = nullptr;
            if (cudaIpcOpenMemHandle(&q, cards[side ? s_rank + 1 : s_rank - 1].snap, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; break; }
            nls_peer_snap[side] = static_cast<char*>(q);
        }
        rc = allgather();   // did every rank manage to map what it needs?
        if (rc) return rc;
        for (const Card& c : cards) all_ok = std::min(all_ok, c.ok);
    }
    cudaFree(d_cards);
    if (all_ok) {
        snap_a.release(); snap_b.release();   // from here on the two arrays live in the IPC block
        snap_a.d = reinterpret_cast<float4*>(nls_snap_block); snap_a.cap = nls_snap_cap;
        snap_b.d = snap_a.d + nls_snap_cap; snap_b.cap = nls_snap_cap;
        nls_ready = true;
    }
    return BLOBS_OK;
}

int World::strip_exchange() {
    if (s_nranks <= 1) return BLOBS_OK;
    if (p2p_on) {
        // One kernel instead of the NCCL group: copies the USED part of both outgoing messages straight into the neighbours'
        // receive buffers (peer stores over NVLink), publishes header + sequence number once every CTA's stores are fenced,
        // and waits for the two incoming sequence numbers. Buffers alternate with the exchange parity: a neighbour cannot be
        // two exchanges ahead of me (it waits for my message of the exchange in between), so parity p is free again.
        const size_t par = (size_t)((nccl_exchanges + 1) & 1u);   // the device-side sequence number (d_push_done[1]) counts in step with nccl_exchanges
        cur_recv[0] = p2p_block + (0 * 2 + par) * p2p_stride;
        cur_recv[1] = p2p_block + (1 * 2 + par) * p2p_stride;
        void* peer_l = strip.has_left ? p2p_peer[0] + (1 * 2 + par) * p2p_stride : nullptr;    // I am my left neighbour's RIGHT
        void* peer_r = strip.has_right ? p2p_peer[1] + (0 * 2 + par) * p2p_stride : nullptr;   // and my right neighbour's LEFT
        BLOBS_LAUNCH(STRIP_PUSH_CTAS, 256, 0, stream, k_strip_push)(strip, msg[0], msg[1], peer_l, peer_r, cur_recv[0], cur_recv[1], d_push_done + 1, d_push_done, d_stats);
        launches++;
        CU(cudaGetLastError());
        nccl_exchanges++;
        return BLOBS_OK;
    }
    ncclComm_t comm = static_cast<ncclComm_t>(nccl_comm);
    NC(g_nccl.GroupStart());
    if (strip.has_right) {
        NC(g_nccl.Send(msg[1], msg_bytes, ncclInt8, s_rank + 1, comm, stream));
        NC(g_nccl.Recv(msg[3], msg_bytes, ncclInt8, s_rank + 1, comm, stream));
    }
    if (strip.has_left) {
        NC(g_nccl.Send(msg[0], msg_bytes, ncclInt8, s_rank - 1, comm, stream));
        NC(g_nccl.Recv(msg[2], msg_bytes, ncclInt8, s_rank - 1, comm, stream));
    }
    NC(g_nccl.GroupEnd());
    nccl_exchanges++;
    return BLOBS_OK;
}

// Everything after the owned colliders were binned into tab_next: [ghost exchange + ghost binning] -> scan -> scatter
// [-> ghost scatter -> ownership hand-over]. Shared by the per-substep path and the out-of-step rebuild.
int World::strip_build_tail(uint32_t* tab_next, uint32_t* tab_cur, uint32_t* tile_next, uint32_t* tile_cur, float4* hot_next, bool timed_launch) {
    const size_t tn = table_entries();
    const uint32_t nc = (uint32_t)cols.slots();
    const BodyArrays B = body_arrays();
    const ColliderArrays C = col_arrays();
    int rc;
    auto run = [&](KClass k, auto&& f) -> int {
        if (timed_launch) return timed(k, f);
        f();
        launches++;
        CU(cudaGetLastError());
        return BLOBS_OK;
    };
    if (strip_on) {
        if (!timed_launch) {  // in-step: the headers are cleared before k_main (launch_substep)
            CU(cudaMemsetAsync(msg[0], 0, sizeof(StripHeader), stream));
            CU(cudaMemsetAsync(msg[1], 0, sizeof(StripHeader), stream));
        }
        if (nc && !timed_launch) {  // inside a step the pack is fused into k_main's tail
            rc = run(KC_PACK, [&] { BLOBS_LAUNCH(cdiv(nc, 256), 256, 0, stream, k_strip_pack)(B, C, strip, d_cowned.d, msg[0], msg[1], nc); });
            if (rc) return rc;
        }
        if (timed_launch && profiling) {
            int rce = BLOBS_OK;
            rc = timed(KC_NCCL, [&] { rce = strip_exchange(); });
            launches--;  // not one of our kernels
            if (rc) return rc;
            if (rce) return rce;
        } else {
            rc = strip_exchange();
            if (rc) return rc;
        }
        rc = run(KC_GHOST, [&] { BLOBS_LAUNCH(cdiv(2 * (size_t)strip.gcap, 256), 256, 0, stream, k_strip_bin_ghosts)(grid, strip, cur_recv[0], cur_recv[1], tab_next, tile_next, gcell.d, d_stats); });
        if (rc) return rc;
    }
    rc = run(KC_SCAN, [&] { BLOBS_LAUNCH(cdiv(tn, SCAN_TILE), SCAN_THREADS, 0, stream, k_scan)(tab_next, (uint32_t)tn, tab_cur, (uint32_t)tn, tile_next, tile_cur); });
    if (rc) return rc;
    if (nc && !strip_on) {
        rc = run(KC_SCATTER, [&] { BLOBS_LAUNCH(cdiv(nc, 256), 256, 0, stream, k_scatter)(C, tab_next, hot_next, nc, nullptr); });
        if (rc) return rc;
    }
    if (strip_on) {
        rc = run(KC_SCATTER, [&] { BLOBS_LAUNCH(cdiv(std::max<uint32_t>(olaunch_dim, 1), 256), 256, 0, stream, k_scatter_owned)(B, C, tab_next, hot_next, olist.d, d_ocount); });
        if (rc) return rc;
        rc = run(KC_GHOST, [&] {
            BLOBS_LAUNCH(cdiv(2 * (size_t)strip.gcap + 4 * (size_t)strip.mcap, 256), 256, 0, stream, k_strip_finish)(B, C, strip, msg[0], msg[1], cur_recv[0], cur_recv[1], tab_next, gcell.d, hot_next,
                                                                                                           d_owned.d, d_cowned.d, olist.d, d_ocount, opos.d, (uint32_t)olist.cap, d_stats);
        });
        if (rc) return rc;
    }
    return BLOBS_OK;
}

int World::strip_owned(uint8_t* out, size_t cap) {
    const size_t n = std::min(cap, bodies.slots());
    if (!strip_on) { std::memset(out, 1, n); return BLOBS_OK; }
    if (n) {
        CU(cudaMemcpyAsync(out, d_owned.d, n, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
    }
    return BLOBS_OK;
}

int World::read_owned_positions(uint32_t* slots, float* xy, size_t cap, size_t* n) {
    int rc = flush();
    if (rc) return rc;
    const size_t nb = bodies.slots();
    if (!d_io_count) CU(cudaMalloc(&d_io_count, sizeof(unsigned int)));
    CU(io_slots.ensure(std::max<size_t>(cap, 1), stream));
    CU(io_xy.ensure(std::max<size_t>(cap, 1), stream));
    CU(cudaMemsetAsync(d_io_count, 0, sizeof(unsigned int), stream));
    if (nb) {
        BLOBS_LAUNCH(cdiv(nb, 256), 256, 0, stream, k_compact_owned)(body_arrays(), strip_on ? d_owned.d : nullptr, (uint32_t)nb, (uint32_t)std::min<size_t>(cap, 0xffffffffu),
                                                           d_io_count, io_slots.d, io_xy.d);
        launches++;
        CU(cudaGetLastError());
    }
    unsigned int cnt = 0;
    CU(cudaMemcpyAsync(&cnt, d_io_count, sizeof(cnt), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    const size_t m = std::min<size_t>(cnt, cap);
    if (m) {
        CU(cudaMemcpyAsync(slots, io_slots.d, m * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        CU(cudaMemcpyAsync(xy, io_xy.d, m * sizeof(float2), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
    }
    if (n) *n = cnt;
    return BLOBS_OK;
}

int World::apply_forces_indexed(const uint32_t* slots, const float* fxy, size_t n) {
    int rc = flush();
    if (rc) return rc;
    if (!n) return BLOBS_OK;
    CU(io_slots.ensure(n, stream));
    CU(d_forces.ensure(n, stream));
    CU(cudaMemcpyAsync(io_slots.d, slots, n * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    CU(cudaMemcpyAsync(d_forces.d, fxy, n * sizeof(float2), cudaMemcpyHostToDevice, stream));
    BLOBS_LAUNCH(cdiv(n, 256), 256, 0, stream, k_apply_forces_indexed)(body_arrays(), strip_on ? d_owned.d : nullptr, io_slots.d, d_forces.d, (uint32_t)n, (uint32_t)bodies.slots());
    launches++;
    CU(cudaGetLastError());
    return BLOBS_OK;
}

// ---------------------------------------------------------------------------------------------- introspection
int World::kernel_info(BlobsKernelInfo* out) const {
    out->launches = launches;
    out->grid_w = grid.W;
    out->grid_h = grid.H;
    out->broadphase_cell = grid.cell;
    out->r_max = r_max;
    out->fused_path = last_fused ? 1u : 0u;
    out->n_simple_bodies = n_simple;
    out->n_multi_bodies = n_multi;
    out->n_spring_bodies = n_sb;
    out->n_islands = n_islands;
    return BLOBS_OK;
}

int World::profile_enable(int on) {
    CU(cudaStreamSynchronize(stream));
    profiling = on != 0;
    profile_main_only = on == 2;   // 2: events around the dominant kernel only (its roofline number, with the rest of the step unperturbed)
    ev_used = 0;
    for (int i = 0; i < KC_COUNT; ++i) { prof_ms[i] = 0.f; prof_launches[i] = 0; }
    return BLOBS_OK;
}

int World::profile_read(float* ms, uint64_t* nl, size_t n) {
    CU(cudaStreamSynchronize(stream));
    int rc = collect_profile();
    if (rc) return rc;
    for (size_t i = 0; i < n && i < (size_t)KC_COUNT; ++i) {
        if (ms) ms[i] = prof_ms[i];
        if (nl) nl[i] = prof_launches[i];
    }
    return BLOBS_OK;
}

}  // namespace blobs
