// extern "C" surface of libblobs_b200.so (include/blobs_b200.h). Thin forwarding only.
#include <cmath>
#include <new>

#include "perf.hpp"
#include "world.hpp"

using blobs::World;

struct BlobsWorld {
    World w;
    explicit BlobsWorld(const BlobsParams& p) : w(p) {}
};

#define W_OR_INVALID(p) \
    if (!(p)) return BLOBS_ERR_INVALID

extern "C" {

int32_t blobs_abi_version(void) { return BLOBS_ABI_VERSION; }

static thread_local char g_create_error[512] = "";

int32_t blobs_world_create(const BlobsParams* params, BlobsWorld** out) {
    if (!params || !out) return BLOBS_ERR_INVALID;
    BlobsWorld* w = new (std::nothrow) BlobsWorld(*params);
    if (!w) return BLOBS_ERR_CAPACITY;
    const int rc = w->w.init();
    if (rc) {
        snprintf(g_create_error, sizeof(g_create_error), "%s", w->w.last_error());
        delete w;
        *out = nullptr;
        return rc;
    }
    *out = w;
    return BLOBS_OK;
}
int32_t blobs_world_destroy(BlobsWorld* w) {
    delete w;
    return BLOBS_OK;
}
int32_t blobs_world_reset(BlobsWorld* w) { W_OR_INVALID(w); return w->w.reset(); }
const char* blobs_last_error(const BlobsWorld* w) { return w ? w->w.last_error() : g_create_error; }
int32_t blobs_world_set_param(BlobsWorld* w, int32_t id, double v) { W_OR_INVALID(w); return w->w.set_param(id, v); }
int32_t blobs_world_get_param(const BlobsWorld* w, int32_t id, double* out) { W_OR_INVALID(w); W_OR_INVALID(out); return w->w.get_param(id, out); }

int32_t blobs_body_insert(BlobsWorld* w, const BlobsBodyDesc* d, BlobsHandle* out) { W_OR_INVALID(w); W_OR_INVALID(d); return w->w.body_insert(*d, out); }
int32_t blobs_body_insert_many(BlobsWorld* w, size_t n, const BlobsBodyDesc* d, BlobsHandle* out) {
    W_OR_INVALID(w);
    if (n && !d) return BLOBS_ERR_INVALID;
    for (size_t i = 0; i < n; ++i) {
        const int rc = w->w.body_insert(d[i], out ? out + i : nullptr);
        if (rc) return rc;
    }
    return BLOBS_OK;
}
int32_t blobs_body_remove(BlobsWorld* w, BlobsHandle h) { W_OR_INVALID(w); return w->w.body_remove(h); }
int32_t blobs_body_get(BlobsWorld* w, BlobsHandle h, BlobsBodyState* out) { W_OR_INVALID(w); W_OR_INVALID(out); return w->w.body_get(h, out); }
int32_t blobs_body_set(BlobsWorld* w, BlobsHandle h, const BlobsBodyState* s, uint32_t mask) { W_OR_INVALID(w); W_OR_INVALID(s); return w->w.body_set(h, *s, mask); }
int32_t blobs_body_count(const BlobsWorld* w, uint64_t* out) { W_OR_INVALID(w); *out = w->w.body_count(); return BLOBS_OK; }
int32_t blobs_body_translate(BlobsWorld* w, BlobsHandle h, BlobsVec2 off) { W_OR_INVALID(w); return w->w.body_translate(h, off); }
int32_t blobs_body_apply_force(BlobsWorld* w, BlobsHandle h, BlobsVec2 f) { W_OR_INVALID(w); return w->w.body_apply_force(h, f); }
int32_t blobs_body_colliders(const BlobsWorld* w, BlobsHandle h, BlobsHandle* out, size_t cap, size_t* n) { W_OR_INVALID(w); W_OR_INVALID(n); return w->w.body_colliders(h, out, cap, n); }

int32_t blobs_collider_insert(BlobsWorld* w, const BlobsColliderDesc* d, BlobsHandle parent, BlobsHandle* out) { W_OR_INVALID(w); W_OR_INVALID(d); return w->w.collider_insert(*d, parent, out); }
int32_t blobs_collider_insert_many(BlobsWorld* w, size_t n, const BlobsColliderDesc* d, const BlobsHandle* parents, BlobsHandle* out) {
    W_OR_INVALID(w);
    if (n && (!d || !parents)) return BLOBS_ERR_INVALID;
    for (size_t i = 0; i < n; ++i) {
        const int rc = w->w.collider_insert(d[i], parents[i], out ? out + i : nullptr);
        if (rc) return rc;
    }
    return BLOBS_OK;
}
int32_t blobs_collider_remove(BlobsWorld* w, BlobsHandle h) { W_OR_INVALID(w); return w->w.collider_remove(h); }
int32_t blobs_collider_get(BlobsWorld* w, BlobsHandle h, BlobsColliderState* out) { W_OR_INVALID(w); W_OR_INVALID(out); return w->w.collider_get(h, out); }
int32_t blobs_collider_count(const BlobsWorld* w, uint64_t* out) { W_OR_INVALID(w); *out = w->w.collider_count(); return BLOBS_OK; }

int32_t blobs_spring_insert(BlobsWorld* w, BlobsHandle a, BlobsHandle b, float rest, float k, float c, BlobsHandle* out) { W_OR_INVALID(w); return w->w.spring_insert(a, b, rest, k, c, out); }
int32_t blobs_spring_remove(BlobsWorld* w, BlobsHandle h) { W_OR_INVALID(w); return w->w.spring_remove(h); }
int32_t blobs_joint_insert(BlobsWorld* w, BlobsHandle a, BlobsHandle b, BlobsVec2 aa, BlobsVec2 ab, float dist, BlobsHandle* out) { W_OR_INVALID(w); return w->w.joint_insert(a, b, aa, ab, dist, out); }
int32_t blobs_spring_insert_many(BlobsWorld* w, size_t n, const BlobsHandle* a, const BlobsHandle* b, const float* p3, BlobsHandle* out) {
    W_OR_INVALID(w);
    if (n && (!a || !b || !p3)) return BLOBS_ERR_INVALID;
    for (size_t i = 0; i < n; ++i) {
        const int rc = w->w.spring_insert(a[i], b[i], p3[3 * i], p3[3 * i + 1], p3[3 * i + 2], out ? out + i : nullptr);
        if (rc) return rc;
    }
    return BLOBS_OK;
}
int32_t blobs_joint_insert_many(BlobsWorld* w, size_t n, const BlobsHandle* a, const BlobsHandle* b, const float* anc, const float* dist, BlobsHandle* out) {
    W_OR_INVALID(w);
    if (n && (!a || !b)) return BLOBS_ERR_INVALID;
    for (size_t i = 0; i < n; ++i) {
        const BlobsVec2 aa = anc ? BlobsVec2{anc[4 * i], anc[4 * i + 1]} : BlobsVec2{0.f, 0.f};
        const BlobsVec2 ab = anc ? BlobsVec2{anc[4 * i + 2], anc[4 * i + 3]} : BlobsVec2{0.f, 0.f};
        const int rc = w->w.joint_insert(a[i], b[i], aa, ab, dist ? dist[i] : NAN, out ? out + i : nullptr);
        if (rc) return rc;
    }
    return BLOBS_OK;
}
int32_t blobs_joint_remove(BlobsWorld* w, BlobsHandle h) { W_OR_INVALID(w); return w->w.joint_remove(h); }
int32_t blobs_constraint_push(BlobsWorld* w, BlobsVec2 p, float r) { W_OR_INVALID(w); return w->w.constraint_push(p, r); }
int32_t blobs_constraint_clear(BlobsWorld* w) { W_OR_INVALID(w); return w->w.constraint_clear(); }

int32_t blobs_step(BlobsWorld* w, double delta, BlobsStepStats* stats) { W_OR_INVALID(w); return w->w.step(delta, 1, stats); }
int32_t blobs_fixed_step(BlobsWorld* w, double frame_time, BlobsStepStats* stats) { W_OR_INVALID(w); return w->w.fixed_step(frame_time, stats); }
int32_t blobs_step_n(BlobsWorld* w, double delta, uint32_t n, BlobsStepStats* stats) { W_OR_INVALID(w); return w->w.step(delta, n, stats); }

int32_t blobs_body_slots(const BlobsWorld* w, uint64_t* out) { W_OR_INVALID(w); *out = w->w.body_slots(); return BLOBS_OK; }
int32_t blobs_collider_slots(const BlobsWorld* w, uint64_t* out) { W_OR_INVALID(w); *out = w->w.collider_slots(); return BLOBS_OK; }
int32_t blobs_download_bodies(BlobsWorld* w, BlobsBodyState* st, BlobsHandle* h, size_t cap) { W_OR_INVALID(w); return w->w.download_bodies(st, h, cap); }
int32_t blobs_download_colliders(BlobsWorld* w, BlobsColliderState* st, BlobsHandle* h, size_t cap) { W_OR_INVALID(w); return w->w.download_colliders(st, h, cap); }
int32_t blobs_read_body_positions(BlobsWorld* w, float* xy, size_t cap) { W_OR_INVALID(w); W_OR_INVALID(xy); return w->w.read_body_vec(0, xy, cap); }
int32_t blobs_read_body_velocities(BlobsWorld* w, float* xy, size_t cap) { W_OR_INVALID(w); W_OR_INVALID(xy); return w->w.read_body_vec(1, xy, cap); }
int32_t blobs_apply_forces(BlobsWorld* w, const float* f, size_t cap) { W_OR_INVALID(w); W_OR_INVALID(f); return w->w.apply_forces(f, cap); }
int32_t blobs_forces_upload_async(BlobsWorld* w, const float* f, size_t cap) { W_OR_INVALID(w); W_OR_INVALID(f); return w->w.forces_upload_async(f, cap); }
int32_t blobs_apply_forces_uploaded(BlobsWorld* w) { W_OR_INVALID(w); return w->w.apply_forces_uploaded(); }
int32_t blobs_read_body_positions_async(BlobsWorld* w, float* xy, size_t cap) { W_OR_INVALID(w); W_OR_INVALID(xy); return w->w.read_positions_async(xy, cap); }
int32_t blobs_io_sync(BlobsWorld* w) { W_OR_INVALID(w); return w->w.io_sync(); }
int32_t blobs_download_cell_coords(BlobsWorld* w, int32_t* cx, int32_t* cy, size_t cap) { W_OR_INVALID(w); W_OR_INVALID(cx); W_OR_INVALID(cy); return w->w.download_cell_coords(cx, cy, cap); }

int32_t blobs_query_circles(BlobsWorld* w, size_t n, const float* c, const float* r, const BlobsQueryFilter* f, uint64_t* off, BlobsHandle* hits, size_t cap, size_t* nh) {
    W_OR_INVALID(w); W_OR_INVALID(off);
    if (n) { W_OR_INVALID(c); W_OR_INVALID(r); }
    if (cap) W_OR_INVALID(hits);
    return w->w.query_circles(n, c, r, f, off, hits, cap, nh);
}
int32_t blobs_debug_counts(const BlobsWorld* w, BlobsDebugCounts* out) { W_OR_INVALID(w); W_OR_INVALID(out); return w->w.debug_counts(out); }
int32_t blobs_debug_data(BlobsWorld* w, float* body_xform, float* joint_ab, float* col_xform, float* col_radius, float* spring_ab, const BlobsDebugCounts* caps) {
    W_OR_INVALID(w); W_OR_INVALID(caps);
    return w->w.debug_data(body_xform, joint_ab, col_xform, col_radius, spring_ab, caps);
}

int32_t blobs_record_contacts(BlobsWorld* w, int32_t mode, size_t cap) { W_OR_INVALID(w); return w->w.record_contacts(mode, cap); }
int32_t blobs_events_drain(BlobsWorld* w, BlobsCollisionEvent* buf, size_t cap, size_t* n) { W_OR_INVALID(w); return w->w.events_drain(buf, cap, n); }
int32_t blobs_pairs_drain(BlobsWorld* w, uint32_t* a, uint32_t* b, size_t cap, size_t* n, uint64_t* se, size_t se_cap, size_t* n_sub) {
    W_OR_INVALID(w);
    return w->w.pairs_drain(a, b, cap, n, se, se_cap, n_sub);
}

int32_t blobs_kernel_info(const BlobsWorld* w, BlobsKernelInfo* out) { W_OR_INVALID(w); W_OR_INVALID(out); return w->w.kernel_info(out); }
int32_t blobs_profile_enable(BlobsWorld* w, int32_t on) { W_OR_INVALID(w); return w->w.profile_enable(on); }
int32_t blobs_profile_read(BlobsWorld* w, float* ms, uint64_t* launches, size_t n) { W_OR_INVALID(w); return w->w.profile_read(ms, launches, n); }

void blobs_perf_counter(const char* name, uint64_t count) { if (name) blobs::PerfCounters::global().update(name, count); }
void blobs_perf_counter_inc(const char* name, uint64_t inc) { if (name) blobs::PerfCounters::global().inc(name, inc); }
void blobs_perf_counters_new_frame(double delta) { blobs::PerfCounters::global().new_frame(delta); }
void blobs_perf_counters_reset(void) { blobs::PerfCounters::global().reset(); }
int32_t blobs_perf_counter_get(const char* name, uint64_t* count, double* avg) {
    if (!name) return BLOBS_ERR_INVALID;
    const blobs::PerfCounter c = blobs::PerfCounters::global().get(name);
    if (count) *count = c.count;
    if (avg) *avg = c.decayed_average;
    return BLOBS_OK;
}
uint64_t blobs_perf_counter_count(void) { return blobs::PerfCounters::global().size(); }
int32_t blobs_perf_counter_at(uint64_t i, char* name, size_t name_cap, uint64_t* count, double* avg) {
    std::string n;
    blobs::PerfCounter c;
    if (!blobs::PerfCounters::global().at((size_t)i, &n, &c)) return BLOBS_ERR_INVALID;
    if (name) {
        if (n.size() + 1 > name_cap) return BLOBS_ERR_CAPACITY;
        std::memcpy(name, n.c_str(), n.size() + 1);
    }
    if (count) *count = c.count;
    if (avg) *avg = c.decayed_average;
    return BLOBS_OK;
}

uint64_t blobs_event_history_len(void) { return blobs::EventHistory::global().size(); }
int32_t blobs_event_history_get(uint64_t i, BlobsPhysicsEvent* out) {
    if (!out) return BLOBS_ERR_INVALID;
    blobs::PhysicsEventRec e;
    if (!blobs::EventHistory::global().at((size_t)i, &e)) return BLOBS_ERR_INVALID;
    std::memset(out, 0, sizeof(*out));
    out->real_time = e.real_time;
    out->unpaused_time = e.unpaused_time;
    out->position = BlobsVec2{e.px, e.py};
    out->has_position = e.has_position ? 1 : 0;
    out->severity = e.severity;
    out->col_handle = e.col_handle;
    out->rbd_handle = e.rbd_handle;
    snprintf(out->message, sizeof(out->message), "%s", e.message.c_str());
    return BLOBS_OK;
}
void blobs_event_history_clear(void) { blobs::EventHistory::global().clear(); }

int32_t blobs_strip_unique_id(uint8_t* out128) {
    if (!out128) return BLOBS_ERR_INVALID;
    std::string e;
    const int rc = World::strip_unique_id(out128, &e);
    if (rc) snprintf(g_create_error, sizeof(g_create_error), "%s", e.c_str());
    return rc;
}
int32_t blobs_strip_configure(BlobsWorld* w, int32_t rank, int32_t nranks, float x_lo, float x_hi, const uint8_t* id128, uint32_t gcap, uint32_t mcap) {
    W_OR_INVALID(w);
    if (nranks > 1 && !id128) return BLOBS_ERR_INVALID;
    return w->w.strip_configure(rank, nranks, x_lo, x_hi, id128, gcap, mcap);
}
int32_t blobs_read_owned_positions(BlobsWorld* w, uint32_t* slots, float* xy, size_t cap, size_t* n) { W_OR_INVALID(w); W_OR_INVALID(slots); W_OR_INVALID(xy); return w->w.read_owned_positions(slots, xy, cap, n); }
int32_t blobs_apply_forces_indexed(BlobsWorld* w, const uint32_t* slots, const float* fxy, size_t n) { W_OR_INVALID(w); if (n && (!slots || !fxy)) return BLOBS_ERR_INVALID; return w->w.apply_forces_indexed(slots, fxy, n); }
int32_t blobs_forces_indexed_upload_async(BlobsWorld* w, const uint32_t* slots, const float* fxy, size_t n) { W_OR_INVALID(w); if (n && (!slots || !fxy)) return BLOBS_ERR_INVALID; return w->w.forces_indexed_upload_async(slots, fxy, n); }
int32_t blobs_apply_forces_indexed_uploaded(BlobsWorld* w) { W_OR_INVALID(w); return w->w.apply_forces_indexed_uploaded(); }
int32_t blobs_read_owned_positions_async(BlobsWorld* w, uint32_t* slots, float* xy, uint32_t* n_out, size_t cap) { W_OR_INVALID(w); W_OR_INVALID(slots); W_OR_INVALID(xy); W_OR_INVALID(n_out); return w->w.read_owned_positions_async(slots, xy, n_out, cap); }
int32_t blobs_strip_owned(BlobsWorld* w, uint8_t* out, size_t cap) { W_OR_INVALID(w); W_OR_INVALID(out); return w->w.strip_owned(out, cap); }

}  // extern "C"
