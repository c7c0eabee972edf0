// TEST INFRASTRUCTURE — NOT PRODUCT CODE. ctypes-facing C ABI over the CPU oracle.
// It re-uses the *struct layouts* of include/blobs_b200.h so tests can feed the same descriptors to
// the oracle and to the CUDA library; it shares no code with the product.
#include <chrono>
#include <cstring>
#include <string>

#include "../include/blobs_b200.h"
#include "blobs_oracle.hpp"

using namespace oracle;

struct OrcWorld {
    Physics p;
    std::string err;
};

static Vec2 V(BlobsVec2 v) { return Vec2{v.x, v.y}; }
static BlobsVec2 B(Vec2 v) { return BlobsVec2{v.x, v.y}; }
static Affine2 A(const BlobsAffine2& a) { return Affine2{Mat2{V(a.x_axis), V(a.y_axis)}, V(a.translation)}; }
static BlobsAffine2 BA(const Affine2& a) { return BlobsAffine2{B(a.matrix2.x_axis), B(a.matrix2.y_axis), B(a.translation)}; }

#define ORC_TRY(w, body)                 \
    try {                                \
        body;                            \
        return 0;                        \
    } catch (const std::exception& e) {  \
        (w)->err = e.what();             \
        return 1;                        \
    }

extern "C" {

OrcWorld* orc_new(float gx, float gy, int use_spatial_hash) {
    auto* w = new OrcWorld();
    w->p.gravity = v2(gx, gy);
    w->p.use_spatial_hash = use_spatial_hash != 0;
    return w;
}
void orc_free(OrcWorld* w) { delete w; }
const char* orc_last_error(OrcWorld* w) { return w->err.c_str(); }

int orc_set_param(OrcWorld* w, int id, double v) {
    Physics& p = w->p;
    switch (id) {
        case BLOBS_PARAM_GRAVITY_X: p.gravity.x = (float)v; break;
        case BLOBS_PARAM_GRAVITY_Y: p.gravity.y = (float)v; break;
        case BLOBS_PARAM_SUBSTEPS: p.substeps = (uint32_t)v; break;
        case BLOBS_PARAM_JOINT_ITERATIONS: p.joint_iterations = (uint32_t)v; break;
        case BLOBS_PARAM_USE_SPATIAL_HASH: p.use_spatial_hash = v != 0; break;
        case BLOBS_PARAM_COLLISIONS_ENABLED: p.collisions_enabled = v != 0; break;
        case BLOBS_PARAM_ACCUMULATOR: p.accumulator = v; break;
        case BLOBS_PARAM_TIME: p.time = v; break;
        case BLOBS_PARAM_OLD_DT: p.old_dt = (float)v; break;
        case BLOBS_PARAM_CELL_SIZE: p.spatial_hash.cell_size = (float)v; break;
        case 100: p.maintain_spatial_hash = v != 0; break;  // oracle-only knobs
        case 101: p.record_events = v != 0; break;
        case 102: p.use_grid_pairs = v != 0; break;
        default: return 1;
    }
    return 0;
}
double orc_get_param(OrcWorld* w, int id) {
    Physics& p = w->p;
    switch (id) {
        case BLOBS_PARAM_GRAVITY_X: return p.gravity.x;
        case BLOBS_PARAM_GRAVITY_Y: return p.gravity.y;
        case BLOBS_PARAM_SUBSTEPS: return p.substeps;
        case BLOBS_PARAM_JOINT_ITERATIONS: return p.joint_iterations;
        case BLOBS_PARAM_USE_SPATIAL_HASH: return p.use_spatial_hash;
        case BLOBS_PARAM_COLLISIONS_ENABLED: return p.collisions_enabled;
        case BLOBS_PARAM_ACCUMULATOR: return p.accumulator;
        case BLOBS_PARAM_TIME: return p.time;
        case BLOBS_PARAM_OLD_DT: return p.old_dt;
        case BLOBS_PARAM_CELL_SIZE: return p.spatial_hash.cell_size;
        default: return 0.0;
    }
}

static RigidBody body_from_desc(const BlobsBodyDesc& d) {  // RigidBodyBuilder::build rigid_body.rs:376-400
    RigidBody r;
    r.position = V(d.position);
    r.position_old = V(d.position_old);
    r.gravity_mod = d.gravity_mod;
    r.rotation = d.rotation;
    r.center_of_mass = v2(0.f, 0.f);
    r.calculated_mass = 1.0f;
    r.angular_velocity = 0.0f;
    r.torque = 0.0f;
    r.inertia = 1.0f;
    r.scale = V(d.scale);
    r.acceleration = V(d.acceleration);
    r.has_velocity_request = d.has_velocity_request != 0;
    r.velocity_request = V(d.velocity_request);
    r.calculated_velocity = V(d.calculated_velocity);
    r.user_data_lo = d.user_data_lo;
    r.user_data_hi = d.user_data_hi;
    r.body_type = (BodyType)d.body_type;
    return r;
}
static Collider col_from_desc(const BlobsColliderDesc& d) {
    Collider c;
    c.offset = A(d.offset);
    c.absolute_transform = A(d.absolute_transform);
    c.radius = d.radius;
    c.has_mass_override = d.has_mass_override != 0;
    c.mass_override = d.mass_override;
    c.is_sensor = d.is_sensor != 0;
    c.memberships = d.memberships;
    c.filter = d.filter;
    c.user_data_lo = d.user_data_lo;
    c.user_data_hi = d.user_data_hi;
    return c;
}

int orc_body_insert_many(OrcWorld* w, size_t n, const BlobsBodyDesc* d, uint64_t* out) {
    ORC_TRY(w, for (size_t i = 0; i < n; ++i) {
        Handle h = w->p.insert_rbd(body_from_desc(d[i]));
        if (out) out[i] = h;
    })
}
int orc_collider_insert_many(OrcWorld* w, size_t n, const BlobsColliderDesc* d, const uint64_t* parents, uint64_t* out) {
    ORC_TRY(w, for (size_t i = 0; i < n; ++i) {
        Handle h = w->p.insert_collider_with_parent(col_from_desc(d[i]), parents[i]);
        if (out) out[i] = h;
    })
}
int orc_body_remove(OrcWorld* w, uint64_t h) { ORC_TRY(w, w->p.remove_rbd(h)) }
int orc_collider_remove(OrcWorld* w, uint64_t h) { ORC_TRY(w, w->p.remove_col(h)) }
int orc_reset(OrcWorld* w) { ORC_TRY(w, w->p.reset()) }

int orc_spring_insert(OrcWorld* w, uint64_t a, uint64_t b, float rest, float k, float c, uint64_t* out) {
    Spring s;
    s.a = a;
    s.b = b;
    s.rest_length = rest;
    s.stiffness = k;
    s.damping = c;
    *out = w->p.springs.insert(s);
    return 0;
}
int orc_spring_remove(OrcWorld* w, uint64_t h) { return w->p.springs.remove(h) ? 0 : 1; }
int orc_joint_insert(OrcWorld* w, uint64_t a, uint64_t b, BlobsVec2 aa, BlobsVec2 ab, float dist, uint64_t* out) {
    ORC_TRY(w, *out = std::isnan(dist) ? w->p.create_fixed_joint(a, b, V(aa), V(ab))
                                       : w->p.create_fixed_joint_with_distance(a, b, V(aa), V(ab), dist))
}
int orc_joint_remove(OrcWorld* w, uint64_t h) { return w->p.joints.remove(h) ? 0 : 1; }
int orc_constraint_push(OrcWorld* w, BlobsVec2 pos, float r) {
    w->p.constraints.push_back(Constraint{V(pos), r});
    return 0;
}
int orc_constraint_clear(OrcWorld* w) {
    w->p.constraints.clear();
    return 0;
}

int orc_step(OrcWorld* w, double delta) { ORC_TRY(w, w->p.step(delta)) }
int orc_fixed_step(OrcWorld* w, double frame_time, int* n_steps) { ORC_TRY(w, *n_steps = w->p.fixed_step(frame_time)) }
// n steps; returns wall seconds in *secs (steady_clock) — used by bench.py's cpu_baseline leg
int orc_step_n_timed(OrcWorld* w, double delta, uint32_t n, double* secs) {
    auto t0 = std::chrono::steady_clock::now();
    try {
        for (uint32_t i = 0; i < n; ++i) w->p.step(delta);
    } catch (const std::exception& e) {
        w->err = e.what();
        return 1;
    }
    *secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

uint64_t orc_body_slots(OrcWorld* w) { return w->p.rbd_set.slots(); }
uint64_t orc_collider_slots(OrcWorld* w) { return w->p.col_set.slots(); }
uint64_t orc_body_count(OrcWorld* w) { return w->p.rbd_set.len; }
uint64_t orc_collider_count(OrcWorld* w) { return w->p.col_set.len; }

static void fill_state(const RigidBody& r, BlobsBodyState& s) {
    s.position = B(r.position);
    s.position_old = B(r.position_old);
    s.center_of_mass = B(r.center_of_mass);
    s.scale = B(r.scale);
    s.acceleration = B(r.acceleration);
    s.velocity_request = B(r.velocity_request);
    s.calculated_velocity = B(r.calculated_velocity);
    s.calculated_mass = r.calculated_mass;
    s.gravity_mod = r.gravity_mod;
    s.rotation = r.rotation;
    s.angular_velocity = r.angular_velocity;
    s.torque = r.torque;
    s.inertia = r.inertia;
    s.has_velocity_request = r.has_velocity_request;
    s.body_type = r.body_type;
    s.user_data_lo = r.user_data_lo;
    s.user_data_hi = r.user_data_hi;
}

int orc_download_bodies(OrcWorld* w, BlobsBodyState* st, uint64_t* handles, size_t cap) {
    auto& a = w->p.rbd_set;
    for (size_t s = 0; s < cap && s < a.slots(); ++s) {
        bool alive = a.alive((uint32_t)s);
        if (handles) handles[s] = alive ? a.handle_at((uint32_t)s) : 0;
        if (st) {
            std::memset(&st[s], 0, sizeof(BlobsBodyState));
            if (alive) fill_state(a.storage[s].value, st[s]);
        }
    }
    return 0;
}
int orc_download_colliders(OrcWorld* w, BlobsColliderState* st, uint64_t* handles, size_t cap) {
    auto& a = w->p.col_set;
    for (size_t s = 0; s < cap && s < a.slots(); ++s) {
        bool alive = a.alive((uint32_t)s);
        if (handles) handles[s] = alive ? a.handle_at((uint32_t)s) : 0;
        if (st) {
            std::memset(&st[s], 0, sizeof(BlobsColliderState));
            if (alive) {
                const Collider& c = a.storage[s].value;
                BlobsColliderDesc& d = st[s].desc;
                d.offset = BA(c.offset);
                d.absolute_transform = BA(c.absolute_transform);
                d.radius = c.radius;
                d.mass_override = c.mass_override;
                d.shape_radius = c.radius;
                d.has_mass_override = c.has_mass_override;
                d.is_sensor = c.is_sensor;
                d.memberships = c.memberships;
                d.filter = c.filter;
                d.user_data_lo = c.user_data_lo;
                d.user_data_hi = c.user_data_hi;
                st[s].parent = c.parent;
            }
        }
    }
    return 0;
}
int orc_body_get(OrcWorld* w, uint64_t h, BlobsBodyState* out) {
    const RigidBody* r = w->p.rbd_set.get(h);
    if (!r) return 1;
    fill_state(*r, *out);
    return 0;
}
int orc_body_set(OrcWorld* w, uint64_t h, const BlobsBodyState* s, uint32_t mask) {
    RigidBody* r = w->p.rbd_set.get(h);
    if (!r) return 1;
    if (mask & BLOBS_BODY_POSITION) r->position = V(s->position);
    if (mask & BLOBS_BODY_POSITION_OLD) r->position_old = V(s->position_old);
    if (mask & BLOBS_BODY_ACCELERATION) r->acceleration = V(s->acceleration);
    if (mask & BLOBS_BODY_VELOCITY_REQUEST) {
        r->has_velocity_request = s->has_velocity_request != 0;
        r->velocity_request = V(s->velocity_request);
    }
    if (mask & BLOBS_BODY_CALC_VELOCITY) r->calculated_velocity = V(s->calculated_velocity);
    if (mask & BLOBS_BODY_ROTATION) r->rotation = s->rotation;
    if (mask & BLOBS_BODY_ANGULAR_VELOCITY) r->angular_velocity = s->angular_velocity;
    if (mask & BLOBS_BODY_TORQUE) r->torque = s->torque;
    if (mask & BLOBS_BODY_MASS) r->calculated_mass = s->calculated_mass;
    if (mask & BLOBS_BODY_INERTIA) r->inertia = s->inertia;
    if (mask & BLOBS_BODY_GRAVITY_MOD) r->gravity_mod = s->gravity_mod;
    if (mask & BLOBS_BODY_TYPE) r->body_type = (BodyType)s->body_type;
    if (mask & BLOBS_BODY_USER_DATA) {
        r->user_data_lo = s->user_data_lo;
        r->user_data_hi = s->user_data_hi;
    }
    if (mask & BLOBS_BODY_SCALE) r->scale = V(s->scale);
    if (mask & BLOBS_BODY_CENTER_OF_MASS) r->center_of_mass = V(s->center_of_mass);
    return 0;
}
// per-slot RigidBody::apply_force (rigid_body.rs:155-160)
int orc_apply_forces(OrcWorld* w, const float* f, size_t cap) {
    auto& a = w->p.rbd_set;
    for (size_t s = 0; s < cap && s < a.slots(); ++s) {
        if (!a.alive((uint32_t)s)) continue;
        RigidBody& r = a.storage[s].value;
        if (!r.is_static()) r.acceleration += v2(f[2 * s], f[2 * s + 1]) / r.calculated_mass;
    }
    return 0;
}
// update_rigid_body_position physics.rs:174-182
int orc_body_translate(OrcWorld* w, uint64_t h, BlobsVec2 off) {
    RigidBody* r = w->p.rbd_set.get(h);
    if (!r) return 1;
    if (w->p.maintain_spatial_hash) w->p.spatial_hash.move_point(h, V(off));
    r->position += V(off);
    return 0;
}
size_t orc_body_colliders(OrcWorld* w, uint64_t h, uint64_t* out, size_t cap) {
    const RigidBody* r = w->p.rbd_set.get(h);
    if (!r) return 0;
    for (size_t i = 0; i < r->colliders.size() && i < cap; ++i) out[i] = r->colliders[i];
    return r->colliders.size();
}

// SpatialHash::get_cell_coords of every collider snapshot with spatial_hash.cell_size
int orc_download_cell_coords(OrcWorld* w, int32_t* cx, int32_t* cy, size_t cap) {
    auto& a = w->p.col_set;
    for (size_t s = 0; s < cap && s < a.slots(); ++s) {
        cx[s] = cy[s] = 0;
        if (a.alive((uint32_t)s)) w->p.spatial_hash.get_cell_coords(a.storage[s].value.absolute_transform.translation, cx[s], cy[s]);
    }
    return 0;
}

uint64_t orc_pairs_count(OrcWorld* w) { return w->p.pair_a.size(); }
uint64_t orc_substeps_recorded(OrcWorld* w) { return w->p.pair_substep_end.size(); }
int orc_pairs_drain(OrcWorld* w, uint32_t* a, uint32_t* b, uint64_t* substep_end) {
    Physics& p = w->p;
    if (a) std::memcpy(a, p.pair_a.data(), p.pair_a.size() * 4);
    if (b) std::memcpy(b, p.pair_b.data(), p.pair_b.size() * 4);
    if (substep_end) std::memcpy(substep_end, p.pair_substep_end.data(), p.pair_substep_end.size() * 8);
    p.pair_a.clear();
    p.pair_b.clear();
    p.pair_substep_end.clear();
    return 0;
}
uint64_t orc_events_count(OrcWorld* w) { return w->p.events.size(); }
int orc_events_drain(OrcWorld* w, BlobsCollisionEvent* out) {
    Physics& p = w->p;
    for (size_t i = 0; i < p.events.size(); ++i) {
        out[i].col_handle_a = p.events[i].col_a;
        out[i].col_handle_b = p.events[i].col_b;
        out[i].impact_vel_a = B(p.events[i].impact_vel_a);
        out[i].impact_vel_b = B(p.events[i].impact_vel_b);
    }
    p.events.clear();
    return 0;
}
uint64_t orc_collisions_total(OrcWorld* w) { return w->p.collisions_total; }
uint64_t orc_coincident_total(OrcWorld* w) { return w->p.coincident_total; }

// ---- standalone SpatialHash (blobs/src/spatial.rs), for the reference's own unit-test vectors
SpatialHash* orc_sh_new(float cs) { return new SpatialHash(cs); }
void orc_sh_free(SpatialHash* s) { delete s; }
uint64_t orc_sh_insert(SpatialHash* s, float x, float y, float r) { return s->insert(v2(x, y), r); }
void orc_sh_insert_with_id(SpatialHash* s, uint64_t id, float x, float y, float r) { s->insert_with_id(id, v2(x, y), r); }
int orc_sh_remove(SpatialHash* s, uint64_t id) { return s->remove(id) ? 1 : 0; }
int orc_sh_move_point(SpatialHash* s, uint64_t id, float dx, float dy) { return s->move_point(id, v2(dx, dy)) ? 1 : 0; }
uint64_t orc_sh_next_id(SpatialHash* s) { return s->next_id; }
void orc_sh_cell_coords(SpatialHash* s, float x, float y, int32_t* cx, int32_t* cy) { s->get_cell_coords(v2(x, y), *cx, *cy); }
size_t orc_sh_cell_population(SpatialHash* s, int32_t cx, int32_t cy) {
    auto it = s->grid.find(SpatialHash::pack(cx, cy));
    return it == s->grid.end() ? 0 : it->second.size();
}
int orc_sh_point(SpatialHash* s, uint64_t id, float* xyr) {
    auto it = s->points.find(id);
    if (it == s->points.end()) return 0;
    xyr[0] = it->second.position.x;
    xyr[1] = it->second.position.y;
    xyr[2] = it->second.radius;
    return 1;
}
size_t orc_sh_query(SpatialHash* s, float x, float y, float r, uint64_t* ids, float* xyr, size_t cap) {
    const auto& res = s->query(v2(x, y), r);
    for (size_t i = 0; i < res.size() && i < cap; ++i) {
        ids[i] = res[i].id;
        xyr[3 * i] = res[i].position.x;
        xyr[3 * i + 1] = res[i].position.y;
        xyr[3 * i + 2] = res[i].radius;
    }
    return res.size();
}

// ---- glam::Affine2 helpers (collider.rs:340-389 vectors)
void orc_body_transform(float rot, float x, float y, BlobsAffine2* out) { *out = BA(affine_from_angle_translation(rot, v2(x, y))); }
void orc_affine_mul(const BlobsAffine2* a, const BlobsAffine2* b, BlobsAffine2* out) { *out = BA(mul(A(*a), A(*b))); }

}  // extern "C"
