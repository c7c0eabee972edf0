// Host side of libblobs_b200: the GPU-resident replacement for the reference's `Physics` struct
// (blobs/src/physics.rs:3-34). Bodies and colliders live in slot-indexed SoA arrays in HBM; the host keeps
// the thunderdome-compatible arenas (slot, generation, LIFO free list) plus the topology (who is whose parent,
// springs, joints) and everything that only changes through the API.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/blobs_b200.h"
#include "types.cuh"

namespace blobs {

// thunderdome::Arena bookkeeping (iteration = ascending slot, LIFO reuse, generation bump on reuse)
struct HostArena {
    std::vector<uint32_t> gen;
    std::vector<uint8_t> alive;
    std::vector<int64_t> next_free;
    int64_t first_free = -1;
    size_t len = 0;

    uint64_t insert() {
        len++;
        if (first_free >= 0) {
            const uint32_t s = (uint32_t)first_free;
            first_free = next_free[s];
            uint32_t g = gen[s] + 1;
            if (g == 0) g = 1;
            gen[s] = g;
            alive[s] = 1;
            return ((uint64_t)g << 32) | s;
        }
        gen.push_back(1);
        alive.push_back(1);
        next_free.push_back(-1);
        return (1ull << 32) | (uint32_t)(gen.size() - 1);
    }
    bool valid(uint64_t h) const {
        const uint32_t s = (uint32_t)h;
        return s < gen.size() && alive[s] && gen[s] == (uint32_t)(h >> 32) && (h >> 32) != 0;
    }
    void remove_slot(uint32_t s) {
        alive[s] = 0;
        next_free[s] = first_free;
        first_free = s;
        len--;
    }
    void clear() {  // Arena::clear == drain
        for (uint32_t s = 0; s < gen.size() && len > 0; ++s)
            if (alive[s]) remove_slot(s);
    }
    size_t slots() const { return gen.size(); }
    uint64_t handle_at(uint32_t s) const { return ((uint64_t)gen[s] << 32) | s; }
};

template <class T>
struct DevBuf {
    T* d = nullptr;
    size_t cap = 0;
    // grows geometrically, preserving contents; new tail is zero-filled
    cudaError_t ensure(size_t n, cudaStream_t st) {
        if (n <= cap) return cudaSuccess;
        size_t ncap = cap ? cap : 256;
        while (ncap < n) ncap *= 2;
        T* nd = nullptr;
        cudaError_t e = cudaMalloc(&nd, ncap * sizeof(T));
        if (e != cudaSuccess) return e;
        if (cap) {
            e = cudaMemcpyAsync(nd, d, cap * sizeof(T), cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) { cudaFree(nd); return e; }
        }
        e = cudaMemsetAsync(nd + cap, 0, (ncap - cap) * sizeof(T), st);
        if (e != cudaSuccess) { cudaFree(nd); return e; }
        if (d) {
            e = cudaStreamSynchronize(st);   // the copy out of the old block must have finished before it is freed
            if (e != cudaSuccess) { cudaFree(nd); return e; }
            cudaFree(d);
        }
        d = nd;
        cap = ncap;
        return cudaSuccess;
    }
    void release() {
        if (d) cudaFree(d);
        d = nullptr;
        cap = 0;
    }
};

// host-authoritative array with a device copy; set() tracks a dirty range, flush() uploads it
template <class T>
struct Mirrored {
    std::vector<T> h;
    DevBuf<T> d;
    size_t lo = SIZE_MAX, hi = 0;
    void resize(size_t n, T fill = T{}) {
        if (n > h.size()) {
            const size_t old = h.size();
            h.resize(n, fill);
            mark(old, n);
        }
    }
    void set(size_t i, const T& v) {
        if (std::memcmp(&h[i], &v, sizeof(T)) != 0) {
            h[i] = v;
            mark(i, i + 1);
        }
    }
    void mark(size_t a, size_t b) {
        if (a < lo) lo = a;
        if (b > hi) hi = b;
    }
    cudaError_t flush(cudaStream_t st) {
        cudaError_t e = d.ensure(h.size(), st);
        if (e != cudaSuccess) return e;
        if (lo < hi) {
            e = cudaMemcpyAsync(d.d + lo, h.data() + lo, (hi - lo) * sizeof(T), cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) return e;
        }
        lo = SIZE_MAX;
        hi = 0;
        return cudaSuccess;
    }
};

struct HBody {
    std::vector<uint64_t> colliders;   // rbd.colliders (each handle twice)
    std::vector<uint64_t> joints;      // rbd.connected_joints
    std::vector<uint32_t> cols;        // distinct live collider slots whose parent is this body
    uint64_t ud_lo = 0, ud_hi = 0;
    BlobsVec2 scale{1.f, 1.f}, com{0.f, 0.f};
    uint32_t type = 0;
    bool rot_active = false;
    uint32_t n_springs = 0, n_joints = 0;
    uint32_t world = 0;               // batched-world id
};

struct HCollider {
    BlobsColliderDesc desc;
    uint64_t parent = 0;
    uint64_t born_epoch = 0;   // World::snap_epoch at insertion: the snapshot matrix is the caller's until a substep has run
};

struct HSpring { uint64_t a, b; float rest, k, c; };
struct HJoint { uint64_t a, b; BlobsVec2 aa, ab; float distance, target; };

enum KClass { KC_MAIN = 0, KC_SCAN, KC_SCATTER, KC_SPRINGS, KC_JOINTS, KC_INTEGRATE, KC_OTHER, KC_PACK, KC_GHOST, KC_NCCL, KC_CROWDED, KC_NLBUILD, KC_DECIDE, KC_COUNT };

class World {
   public:
    explicit World(const BlobsParams& p);
    ~World();
    int init();

    // API (see include/blobs_b200.h)
    int reset();
    int set_param(int id, double v);
    int get_param(int id, double* out) const;
    int body_insert(const BlobsBodyDesc& d, uint64_t* out);
    int body_remove(uint64_t h);
    int body_get(uint64_t h, BlobsBodyState* out);
    int body_set(uint64_t h, const BlobsBodyState& s, uint32_t mask);
    int body_translate(uint64_t h, BlobsVec2 off);
    int body_apply_force(uint64_t h, BlobsVec2 f);
    int body_colliders(uint64_t h, uint64_t* out, size_t cap, size_t* n) const;
    int collider_insert(const BlobsColliderDesc& d, uint64_t parent, uint64_t* out);
    int collider_remove(uint64_t h);
    int collider_get(uint64_t h, BlobsColliderState* out);
    int spring_insert(uint64_t a, uint64_t b, float rest, float k, float c, uint64_t* out);
    int spring_remove(uint64_t h);
    int joint_insert(uint64_t a, uint64_t b, BlobsVec2 aa, BlobsVec2 ab, float dist, uint64_t* out);
    int joint_remove(uint64_t h);
    int constraint_push(BlobsVec2 p, float r);
    int constraint_clear();
    int step(double delta, uint32_t n, BlobsStepStats* stats);
    int fixed_step(double frame_time, BlobsStepStats* stats);
    int download_bodies(BlobsBodyState* st, uint64_t* handles, size_t cap);
    int download_colliders(BlobsColliderState* st, uint64_t* handles, size_t cap);
    int query_circles(size_t n, const float* centre_xy, const float* radius, const BlobsQueryFilter* filter, uint64_t* offsets, uint64_t* hits,
                      size_t hit_cap, size_t* n_hits);
    int debug_counts(BlobsDebugCounts* out) const;
    int debug_data(float* body_xform, float* joint_ab, float* col_xform, float* col_radius, float* spring_ab, const BlobsDebugCounts* caps);
    int read_body_vec(int which, float* xy, size_t cap);
    int apply_forces(const float* f, size_t cap);
    int forces_upload_async(const float* f, size_t cap);
    int apply_forces_uploaded();
    int read_positions_async(float* xy, size_t cap);
    int io_sync();
    int forces_indexed_upload_async(const uint32_t* slots, const float* fxy, size_t n);
    int apply_forces_indexed_uploaded();
    int read_owned_positions_async(uint32_t* slots, float* xy, uint32_t* n_out, size_t cap);
    int download_cell_coords(int32_t* cx, int32_t* cy, size_t cap);
    int record_contacts(int mode, size_t cap);
    int events_drain(BlobsCollisionEvent* buf, size_t cap, size_t* n);
    int pairs_drain(uint32_t* a, uint32_t* b, size_t cap, size_t* n, uint64_t* sub_end, size_t sub_cap, size_t* n_sub);
    int kernel_info(BlobsKernelInfo* out) const;
    int profile_enable(int on);
    int profile_read(float* ms, uint64_t* launches, size_t n);
    static int strip_unique_id(uint8_t* out128, std::string* err);
    int strip_configure(int rank, int nranks, float x_lo, float x_hi, const uint8_t* id128, uint32_t gcap, uint32_t mcap);
    int strip_owned(uint8_t* out, size_t cap);
    int read_owned_positions(uint32_t* slots, float* xy, size_t cap, size_t* n);
    int apply_forces_indexed(const uint32_t* slots, const float* fxy, size_t n);

    size_t body_slots() const { return bodies.slots(); }
    size_t collider_slots() const { return cols.slots(); }
    size_t body_count() const { return bodies.len; }
    size_t collider_count() const { return cols.len; }
    const char* last_error() const { return err.c_str(); }

   private:
    int fail(int code, const std::string& msg) {
        err = msg;
        return code;
    }
    int cuda_fail(cudaError_t e, const char* what);
    int flush();
    int rebuild_topology();
    int rebuild_broadphase();
    int choose_grid(bool force);
    int integrate(uint32_t substeps, float delta, bool last_of_call);
    int run_step(uint32_t substeps, float delta, bool last_of_call, bool allow_graph);
    uint64_t step_key(uint32_t substeps, float delta, bool last_of_call);
    int launch_substep(const SubstepParams& P_in);
    int finish_stats(BlobsStepStats* out, uint32_t steps, uint32_t substeps_run);
    int ensure_shadow();
    int peek_position(uint32_t slot, float2* out);
    void update_mass_and_inertia(uint32_t bslot);
    BodyWrite& stage(uint32_t slot);
    int flush_writes();
    int ensure_capacity();
    BodyArrays body_arrays();
    ColliderArrays col_arrays();
    Constraints constraints_pod() const;
    template <class F>
    int timed(KClass k, F&& f);
    int collect_profile();

    BlobsParams params;
    std::string err;
    int device = 0;
    cudaStream_t stream = nullptr;

    // pub fields of Physics
    float gx = 0.f, gy = 0.f;
    uint32_t substeps = 8, joint_iterations = 4;
    bool use_spatial_hash = false, collisions_enabled = true;
    double accumulator = 0.0, time = 0.0;
    float old_dt = 1.0f;
    float cell_size = 2.0f;         // spatial_hash.cell_size
    float bp_cell_override = 0.0f;  // 0 = auto
    int contact_mode = 0;           // 0 ordered, 1 fast
    bool allow_fused = true;
    int tune = 0;                   // kernel-variant selector (benchmarking aid)
    int crowded_mode = 2;           // BLOBS_PARAM_CROWDED: 0 inline, 1 always defer to k_crowded, 2 auto
    bool crowded_seen = false;      // auto mode: a recent step call reported contact-list overflows
    int pool_mode = 2;              // BLOBS_PARAM_POOL: 0 per-lane contact resolution, 1 warp-pooled k_main, 2 auto (by contact density)
    uint32_t pool_min = 16;         // BLOBS_PARAM_POOL_MIN: survivors per warp from which the pooled path is taken
    bool pool_seen = false;
    int pool_hold = 0;
    int crowded_hold = 0;           //            step calls left before auto mode drops k_crowded again
    std::vector<BlobsVec2> con_pos;
    std::vector<float> con_r;
    bool con_dirty = true;
    DevBuf<float4> d_constraints;

    // arenas + host records
    HostArena bodies, cols, springs, joints;
    uint64_t snap_epoch = 0;   // number of step calls that ran at least one substep (every live collider snapshot is rewritten by each)
    BlobsAffine2 live_snapshot(uint32_t col_slot, float2 translation, const std::vector<float>& rot_host) const;
    std::vector<HBody> hb;
    std::vector<HCollider> hc;
    std::vector<HSpring> hs;
    std::vector<HJoint> hj;

    // host-authoritative device arrays
    Mirrored<float> inertia;
    Mirrored<uint2> binfo;    // (BF_* flags, body_col)
    Mirrored<float2> bmg;     // (calculated_mass, gravity_mod)
    Mirrored<uint32_t> bworld; // batched-world id per body
    uint32_t cur_world = 0, n_worlds = 1;
    size_t table_entries() const { return (size_t)grid.n_worlds * grid.ncells + 1; }
    Mirrored<float2> coff;
    Mirrored<uint4> cconst;   // (radius bits, CF_* flags, memberships, filter)
    Mirrored<uint32_t> cparent;
    Mirrored<uint4> ccold;    // (parent mass bits, memberships, filter, parent slot)
    float mass_of(uint32_t s) const { return bmg.h[s].x; }
    void set_mass(uint32_t s, float m) { float2 e = bmg.h[s]; e.x = m; bmg.set(s, e); }
    void set_gmod(uint32_t s, float g) { float2 e = bmg.h[s]; e.y = g; bmg.set(s, e); }
    // device-authoritative body state
    DevBuf<float2> pos, pos_old, acc, vel, vreq, cabs;
    DevBuf<uint8_t> has_vreq;
    DevBuf<float> rot, angvel, torque;
    DevBuf<uint2> ccell;
    // staged writes
    std::vector<BodyWrite> pending;
    std::vector<int32_t> pending_idx;
    std::vector<ColWrite> pending_col;
    DevBuf<BodyWrite> d_pending;
    DevBuf<ColWrite> d_pending_col;
    // host shadow of pos/rot (valid between a sync point and the next step)
    bool shadow_valid = false;
    std::vector<float2> sh_pos;
    std::vector<float> sh_rot;

    // derived topology
    bool topo_dirty = true, bp_dirty = true;
    int topo_error = 0;
    std::string topo_error_msg;
    bool any_dynamic = false;
    float r_max = 0.f;
    uint32_t n_simple = 0, n_multi = 0, n_sb = 0, n_islands = 0, n_springs_live = 0, n_joints_live = 0, n_active_cols = 0, n_loose = 0;
    DevBuf<uint32_t> mb_body, mb_off, mb_cols, sb_body, sb_off, sb_edge, isl_off, isl_joint;
    DevBuf<SpringParams> d_springs;
    DevBuf<JointParams> d_joints;
    DevBuf<float4> d_joints_inter;        // per-CTA interleaved joint records with island-local body indices (k_joints_fused)
    uint32_t isl_max_joints = 0;
    DevBuf<uint32_t> isl_boff, isl_body;
    uint32_t isl_max_bodies = 0;
    bool joints_smem_ok = false;

    // broadphase
    GridDesc grid{1, 1, 1, 1, 1.0f, 1.0f, 0.f, 0ull, 0ull};
    DevBuf<float4> hot_a, hot_b;
    DevBuf<uint32_t> over_list;        // bodies deferred to k_crowded this substep (capacity = body slots)
    DevBuf<uint32_t> tile_a, tile_b;   // per-scan-tile totals, paired with tab_a / tab_b
    DevBuf<uint32_t> tab_a, tab_b;
    bool cur_is_a = true;
    int bb[4] = {0, 0, 0, 0};
    bool bb_valid = false;

    // neighbour-list pipeline (nlist.cuh): lists of every collider within r_a + r_b + skin, rebuilt on the device's own decision
    int list_mode = 2;                 // BLOBS_PARAM_LIST: 0 = cell grid rebuilt every substep (k_main), 1 = neighbour lists (k_step), 2 = automatic
    int nl_grid_hold = 0, nl_next_hold = 32;   // automatic mode: step calls left on the grid pipeline before lists are tried again
    unsigned long long nl_seen_rebuilds = 0, nl_seen_substeps = 0;
    float skin_frac = 0.8f;            // BLOBS_PARAM_SKIN: skin as a fraction of the largest collider radius
    float nl_skin = 0.f;
    bool nl_on = false;                // the current broadphase is the list pipeline
    bool nl_force_pending = false;     // host-side changes since the last step that invalidate the lists
    bool nl_collisions_were_on = true;
    DevBuf<float4> snap_a, snap_b;     // slot-indexed snapshot records, double-buffered by cur_is_a
    DevBuf<uint4> nl_hdr;
    DevBuf<uint32_t> nl_idx;
    uint32_t nl_stride = 0;
    NlCtl* d_nlctl = nullptr;
    NlCtl* h_nlctl = nullptr;          // pinned copy, refreshed with the step statistics
    // ... on a strip-decomposed world: the snapshot arrays and a flag block are CUDA-IPC memory mapped by the other ranks
    bool nls_ready = false;            // every rank mapped what it needs: lists may be used while strips are on
    char* nls_snap_block = nullptr;    // snap_a | snap_b (then owned by this block, not by the DevBufs)
    char* nls_peer_snap[2] = {nullptr, nullptr};      // left / right neighbour's block
    NlFlag* nls_flags = nullptr;
    NlFlag* nls_peer_flags[NL_MAX_RANKS] = {};
    size_t nls_snap_cap = 0;
    int nls_setup();
    NlStripDev nls_dev();
    NlView nl_view();
    int nl_rebuild_chain(bool timed_launch, bool decide);
    // captured steps: the rebuild kernels of a substep live in the body of a conditional (IF) graph node
    bool cond_nodes = false;           // BLOBS_B200_COND=1: opt-in. Measured on B200 (profiles/r2_notes.md): an IF node costs MORE than the four
                                       // self-gating launches it replaces (0.586 vs 0.546 ms per step on config #2), so it is off by default
    cudaStream_t s_body = nullptr;     // capture stream for the IF-node bodies
    unsigned long long nl_cond_pending = 0ull;   // handle created for the NEXT substep's IF node (set by this substep's k_step)
    uint32_t nl_sub_i = 0, nl_sub_n = 0;
    uint64_t cond_launches = 0, cond_nodes_built = 0, cond_per_rebuild_live = 0;
    bool nl_prev_tail = false;         // the previous substep's k_step already took the rebuild decision for this one
    int nl_rebuild_now();

    // stats / recording
    DeviceStats* d_stats = nullptr;
    DeviceStats* h_stats = nullptr;  // pinned
    int rec_mode = 0;
    size_t rec_cap = 0;
    size_t rec_drained = 0;            // events of the current recording already handed out by partial blobs_events_drain calls
    DevBuf<uint2> rec_pairs;
    DevBuf<float4> rec_vels;
    unsigned long long* d_rec_count = nullptr;
    DevBuf<unsigned long long> d_sub_end;
    std::vector<uint64_t> sub_end_host;  // running pair counts per substep since last drain
    uint32_t sub_recorded = 0;
    bool last_fused = false;

    // strip decomposition (config #5)
    bool strip_on = false;
    int s_rank = 0, s_nranks = 1;
    StripDesc strip{};
    DevBuf<uint8_t> d_owned, d_cowned;
    DevBuf<uint2> gcell;
    DevBuf<uint32_t> olist, opos;      // compact list of owned body slots + position of each body in it
    uint32_t* d_ocount = nullptr;
    uint32_t olaunch = 0;              // host upper bound of *d_ocount for launch sizing
    StripView strip_view();
    int strip_rebuild_olist();
    DevBuf<uint32_t> io_slots;
    DevBuf<float2> io_xy;
    unsigned int* d_io_count = nullptr;
    void* msg[4] = {nullptr, nullptr, nullptr, nullptr};  // send_l, send_r, recv_l, recv_r
    size_t msg_bytes = 0;
    void* nccl_comm = nullptr;
    uint64_t nccl_exchanges = 0;
    uint32_t last_max_ghosts = 0, last_max_migrants = 0;
    int strip_exchange();
    // peer-memory exchange (BLOBS_PARAM_STRIP_P2P): my receive block = 4 messages [from left / from right][exchange parity];
    // the neighbours' blocks are mapped through CUDA IPC and written by k_strip_push over NVLink
    bool p2p_request = false, p2p_on = false;
    char* p2p_block = nullptr;          // mine
    char* p2p_peer[2] = {nullptr, nullptr};   // left / right neighbour's block (IPC mapping)
    size_t p2p_stride = 0;
    unsigned int* d_push_done = nullptr;   // [0] CTA arrival counter, [1] exchange sequence number (device-resident: graph replay)
    bool strip_graph = true;                // with the peer-memory exchange, whole steps are replayed as CUDA graphs (BLOBS_B200_STRIP_GRAPH=0: plain launches)
    bool joint_advance = true;              // k_joints_fused advances the jointed bodies itself instead of writing them back for a k_integrate pass to re-read
                                            // (config #4: 3.31 -> 3.04 ms per step; BLOBS_B200_JADV=0 restores the separate pass)
    bool nls_tail_publish = false;          // BLOBS_B200_NLS_TAIL=1: k_step's last CTA publishes the end-of-substep flags instead of k_nls_publish. Measured on 2x B200
                                            // (profiles/r2_notes.md): SLOWER, 1.646 vs 1.464 ms per step - a gpu-scope fence per CTA waits for that CTA's peer stores
    void* cur_recv[2] = {nullptr, nullptr};   // receive buffers of the exchange in flight (== msg[2], msg[3] on the NCCL path)
    int strip_p2p_setup();
    int strip_build_tail(uint32_t* tab_next, uint32_t* tab_cur, uint32_t* tile_next, uint32_t* tile_cur, float4* hot_next, bool timed_launch);

    // profiling
    uint64_t launches = 0;
    bool profiling = false, profile_main_only = false;
    struct EvPair { cudaEvent_t a, b; int k; };
    // one captured CUDA graph per (last-step-of-call?) flavour of Physics::integrate; replayed while its key matches
    struct GraphSlot { cudaGraphExec_t exec = nullptr; uint64_t key = 0, launches = 0, cond_per_rebuild = 0; std::vector<EvPair> evs; bool profiled = false; };
    GraphSlot gslot[2];
    bool graphs_on = true, capturing = false;
    std::vector<EvPair>* cap_evs = nullptr;
    std::vector<GraphSlot*> graphs_launched;
    uint64_t graph_replays = 0, graph_captures = 0;
    uint32_t olaunch_dim = 0;          // strip mode: launch bound, constant within a blobs_step* call
    void destroy_graph(GraphSlot& g);
    std::vector<EvPair> ev_pool;
    size_t ev_used = 0;
    float prof_ms[KC_COUNT] = {0};
    uint64_t prof_launches[KC_COUNT] = {0};
    cudaEvent_t ev_step0 = nullptr, ev_step1 = nullptr;
    DevBuf<float2> d_forces;
    // pipelined host I/O (blobs_forces_upload_async / blobs_read_body_positions_async): one copy stream per direction so the
    // PCIe transfers of neighbouring steps overlap the kernels of this one
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_up_done[2] = {nullptr, nullptr}, ev_up_free[2] = {nullptr, nullptr}, ev_snap_ready = nullptr, ev_snap_free = nullptr;
    DevBuf<float2> d_forces_up[2], d_pos_snap;
    DevBuf<uint32_t> d_fslots_up[2], d_oslots_snap;   // indexed (strip) forms: slot lists beside the force / position payloads
    DevBuf<float2> d_oxy_snap;
    unsigned int* d_ocount_snap = nullptr;
    bool up_indexed = false;
    int up_next = 0, up_pending = -1;
    size_t up_n = 0;
    bool up_used[2] = {false, false}, snap_used = false, io_ready = false;
    int io_init();
    DevBuf<int> d_cellx, d_celly;
    DevBuf<float2> d_qcentre;
    DevBuf<float> d_qradius;
    DevBuf<uint32_t> d_qcount, d_qoff, d_qhits;
};

}  // namespace blobs
