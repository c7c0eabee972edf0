// TEST INFRASTRUCTURE — NOT PRODUCT CODE. See blobs_oracle.hpp for scope and parity status.
// Every function cites the reference lines (relative to /root/reference) it restates.
#include "blobs_oracle.hpp"

#include <algorithm>
#include <cstdio>

namespace oracle {

// physics.rs:71-76 — clears the four arenas only (not spatial_hash / constraints / time).
void Physics::reset() {
    rbd_set.clear();
    col_set.clear();
    joints.clear();
    springs.clear();
}

// physics.rs:78-82
void Physics::step(double delta) {
    integrate(substeps, (float)delta);
    time += delta;
}

// physics.rs:84-99
int Physics::fixed_step(double frame_time) {
    accumulator += frame_time;
    const double delta = 1.0 / 60.0;
    int max_steps = 3, n = 0;
    while (accumulator >= delta && max_steps > 0) {
        integrate(substeps, (float)delta);
        accumulator -= delta;
        time += delta;
        max_steps -= 1;
        n++;
    }
    return n;
}

// physics.rs:121-128
Handle Physics::insert_rbd(const RigidBody& rbd) {
    Vec2 position = rbd.position;
    Handle h = rbd_set.insert(rbd);
    if (maintain_spatial_hash) spatial_hash.insert_with_id(h, position, 0.5f);
    return h;
}

// rigid_body.rs:96-128
void Physics::update_mass_and_inertia(RigidBody& body) {
    body.calculated_mass = 0.0f;
    body.inertia = 0.0f;
    Vec2 weighted_centers{0.f, 0.f};
    for (Handle ch : body.colliders) {
        const Collider* c = col_set.get(ch);
        if (!c) continue;  // eprintln only
        if (c->is_sensor) continue;
        float mass = c->mass();
        body.calculated_mass += mass;
        body.inertia += c->inertia();
        weighted_centers += c->offset.translation * mass;
    }
    if (body.calculated_mass == 0.0f) body.calculated_mass = 1.0f;
    if (body.inertia == 0.0f) body.inertia = 1.0f;
    body.center_of_mass = weighted_centers / body.calculated_mass;
}

// physics.rs:130-149 + collider.rs:166-196 (handle pushed twice: SURVEY Q1)
Handle Physics::insert_collider_with_parent(Collider col, Handle rbd_handle) {
    col.parent = rbd_handle;                                   // collider.rs:172
    Handle ch = col_set.insert(col);                           // collider.rs:176
    if (RigidBody* r = rbd_set.get(rbd_handle)) r->colliders.push_back(ch);  // collider.rs:179-181
    RigidBody* rbd = rbd_set.get(rbd_handle);
    if (!rbd) throw OraclePanic("parent rigid body must exist when inserting collider");  // physics.rs:142
    rbd->colliders.push_back(ch);   // physics.rs:144
    update_mass_and_inertia(*rbd);  // physics.rs:146
    return ch;
}

// physics.rs:159-161 -> collider.rs:134-164
void Physics::remove_col(Handle h) {
    bool remove_rbd_flag = false;
    if (Collider* c = col_set.get(h)) {
        if (c->parent != NO_HANDLE) {
            Handle parent = c->parent;
            if (RigidBody* body = rbd_set.get(parent)) {
                auto& v = body->colliders;
                v.erase(std::remove(v.begin(), v.end(), h), v.end());  // retain(|&x| x != handle)
                update_mass_and_inertia(*body);  // NB: the collider is still in col_set here, but no longer listed
                if (v.empty()) remove_rbd_flag = true;
            }
            if (remove_rbd_flag) rbd_set.remove(parent);  // rigid_body.rs:265-276 (spatial hash point is NOT removed)
        }
    }
    col_set.remove(h);  // remove_ignoring_parent
}

// physics.rs:163-172
void Physics::remove_rbd(Handle h) {
    if (RigidBody* rbd = rbd_set.get(h)) {
        std::vector<Handle> cols = rbd->colliders;
        for (Handle ch : cols) col_set.remove(ch);
    }
    rbd_set.remove(h);
    if (maintain_spatial_hash) spatial_hash.remove(h);
}

// physics.rs:184-207
Handle Physics::create_fixed_joint(Handle a, Handle b, Vec2 anchor_a, Vec2 anchor_b) {
    if (h_slot(a) == h_slot(b)) throw OraclePanic("get2_mut called with identical indices");
    RigidBody* ra = rbd_set.get(a);
    RigidBody* rb = rbd_set.get(b);
    if (!ra || !rb) throw OraclePanic("create_fixed_joint: unwrap on None");
    float distance = length(ra->position + anchor_a - rb->position - anchor_b);  // physics.rs:198
    return create_fixed_joint_with_distance(a, b, anchor_a, anchor_b, distance);
}

// physics.rs:209-239
Handle Physics::create_fixed_joint_with_distance(Handle a, Handle b, Vec2 anchor_a, Vec2 anchor_b, float distance) {
    if (h_slot(a) == h_slot(b)) throw OraclePanic("get2_mut called with identical indices");
    RigidBody* ra = rbd_set.get(a);
    RigidBody* rb = rbd_set.get(b);
    if (!ra || !rb) throw OraclePanic("create_fixed_joint_with_distance: unwrap on None");
    FixedJoint j;
    j.a = a;
    j.b = b;
    j.anchor_a = anchor_a;
    j.anchor_b = anchor_b;
    j.distance = distance;
    j.target_angle = rb->rotation - ra->rotation;  // physics.rs:230
    Handle jh = joints.insert(j);
    ra->connected_joints.push_back(jh);
    rb->connected_joints.push_back(jh);
    return jh;
}

// physics.rs:369-375 + rigid_body.rs:207-209
void Physics::apply_gravity() {
    for (uint32_t s = 0; s < rbd_set.slots(); ++s) {
        if (!rbd_set.alive(s)) continue;
        RigidBody& body = rbd_set.storage[s].value;
        if (!body.is_static()) body.acceleration += gravity * body.gravity_mod;
    }
}

// springs.rs:25-47 + rigid_body.rs:155-160
void Physics::apply_spring(const Spring& sp) {
    if (h_slot(sp.a) == h_slot(sp.b)) throw OraclePanic("spring: get2_mut identical indices");
    RigidBody* a = rbd_set.get(sp.a);
    RigidBody* b = rbd_set.get(sp.b);
    if (!a || !b) throw OraclePanic("spring: zip_unwrap on None");
    Vec2 delta_position = b->position - a->position;
    float distance = length(delta_position);
    Vec2 direction = delta_position / distance;
    Vec2 relative_velocity = a->calculated_velocity - b->calculated_velocity;
    Vec2 damping_force = sp.damping * dot(relative_velocity, direction) * direction;
    Vec2 force_magnitude = sp.stiffness * (distance - sp.rest_length) - damping_force;  // f32 - Vec2 (Q7)
    Vec2 force = direction * force_magnitude;
    if (!a->is_static()) a->acceleration += force / a->calculated_mass;
    if (!b->is_static()) b->acceleration += (-force) / b->calculated_mass;
}

// The body of the pair loop, physics.rs:250-312, for colliders in slots (slot_a > slot_b in
// iteration terms: a = outer/later key, b = inner/earlier key).
void Physics::resolve_pair(uint32_t slot_a, uint32_t slot_b, uint64_t& count) {
    Collider& col_a = col_set.storage[slot_a].value;
    Collider& col_b = col_set.storage[slot_b].value;
    if (col_a.parent == NO_HANDLE) return;  // physics.rs:252
    if (col_b.parent == NO_HANDLE) return;  // physics.rs:253
    Handle parent_a = col_a.parent, parent_b = col_b.parent;
    // groups.rs:52-57
    if (!((col_a.memberships & col_b.filter) != 0 && (col_b.memberships & col_a.filter) != 0)) return;
    if (parent_a == parent_b) return;  // physics.rs:260

    Vec2 axis = col_a.absolute_transform.translation - col_b.absolute_transform.translation;  // :264
    float distance = length(axis);
    float min_dist = col_a.radius + col_b.radius;  // :267

    if (distance < min_dist) {
        RigidBody* rbd_a = rbd_set.get(parent_a);
        RigidBody* rbd_b = rbd_set.get(parent_b);
        if (!rbd_a || !rbd_b) return;  // :270 (stale parent handle)

        if (distance < 1e-6f) {  // :272-286
            Vec2 push_out = v2(0.01f, 0.0f);
            rbd_a->position += push_out;
            rbd_b->position -= push_out;
            col_a.absolute_transform.translation = rbd_a->position + col_a.offset.translation;
            col_b.absolute_transform.translation = rbd_b->position + col_b.offset.translation;
            axis = col_a.absolute_transform.translation - col_b.absolute_transform.translation;
            distance = length(axis);
            coincident_total++;
        }

        Vec2 impact_vel_a = rbd_a->calculated_velocity;  // :288-289
        Vec2 impact_vel_b = rbd_b->calculated_velocity;

        if (!col_a.is_sensor && !col_b.is_sensor) {  // :291-300
            Vec2 n = axis / distance;
            if (is_nan(n)) throw OraclePanic("assertion failed: !n.is_nan()");
            float delta = min_dist - distance;
            float ratio = 1.0f - rbd_a->calculated_mass / (rbd_a->calculated_mass + rbd_b->calculated_mass);  // :319-321
            rbd_a->position += ratio * delta * n;
            rbd_b->position -= (1.0f - ratio) * delta * n;
        }

        count += 1;
        pair_a.push_back(slot_a);
        pair_b.push_back(slot_b);
        if (record_events)
            events.push_back(CollisionEvent{col_set.handle_at(slot_a), col_set.handle_at(slot_b), impact_vel_a, impact_vel_b});
    }
}

// physics.rs:241-317
void Physics::brute_force_collisions() {
    std::vector<uint32_t> keys;  // :244 — slot order
    keys.reserve(col_set.len);
    for (uint32_t s = 0; s < col_set.slots(); ++s)
        if (col_set.alive(s)) keys.push_back(s);
    uint64_t count = 0;
    for (size_t i = 0; i < keys.size(); ++i)
        for (size_t j = 0; j < i; ++j) resolve_pair(keys[i], keys[j], count);
    collisions_total += count;  // :316
}

// "Grid oracle" (SURVEY §7.1 step 0): NOT the reference algorithm. Candidate pairs come from a
// cell list over the snapshot positions; they are then sorted into the brute-force loop's
// (i, j<i) order and resolved by the very same resolve_pair(), so results are identical to
// brute_force_collisions() as long as no coincident-centre pair (physics.rs:272-286) rewrites a
// snapshot mid-loop (coincident_total is exported so tests can assert that).
void Physics::grid_collisions() {
    std::vector<uint32_t> keys;
    float rmax = 0.f;
    for (uint32_t s = 0; s < col_set.slots(); ++s)
        if (col_set.alive(s)) {
            keys.push_back(s);
            rmax = std::max(rmax, col_set.storage[s].value.radius);
        }
    if (keys.empty()) return;
    const float cs = std::max(2.0f * rmax * 1.001f, 1e-3f);
    std::unordered_map<uint64_t, std::vector<uint32_t>> cells;
    cells.reserve(keys.size() * 2);
    auto cell_of = [&](Vec2 p, int32_t& cx, int32_t& cy) {
        cx = SpatialHash::f2i_sat(std::floor(p.x / cs));
        cy = SpatialHash::f2i_sat(std::floor(p.y / cs));
    };
    for (uint32_t s : keys) {
        int32_t cx, cy;
        cell_of(col_set.storage[s].value.absolute_transform.translation, cx, cy);
        cells[SpatialHash::pack(cx, cy)].push_back(s);
    }
    std::vector<uint64_t> cand;  // (a<<32 | b), a > b
    for (uint32_t s : keys) {
        const Collider& c = col_set.storage[s].value;
        int32_t cx, cy;
        cell_of(c.absolute_transform.translation, cx, cy);
        for (int dx = -2; dx <= 2; ++dx)      // 5x5: one spare ring guards float rounding at cell edges
            for (int dy = -2; dy <= 2; ++dy) {
                auto it = cells.find(SpatialHash::pack(cx + dx, cy + dy));
                if (it == cells.end()) continue;
                for (uint32_t o : it->second) {
                    if (o >= s) continue;
                    const Collider& d = col_set.storage[o].value;
                    Vec2 axis = c.absolute_transform.translation - d.absolute_transform.translation;
                    if (length(axis) < c.radius + d.radius) cand.push_back(((uint64_t)s << 32) | o);
                }
            }
    }
    std::sort(cand.begin(), cand.end());
    uint64_t count = 0;
    for (uint64_t k : cand) resolve_pair((uint32_t)(k >> 32), (uint32_t)(k & 0xffffffffu), count);
    collisions_total += count;
}

// physics.rs:424-477
void Physics::solve_fixed_joints(float dt) {
    for (uint32_t it = 0; it < joint_iterations; ++it) {
        for (uint32_t s = 0; s < joints.slots(); ++s) {
            if (!joints.alive(s)) continue;
            const FixedJoint& joint = joints.storage[s].value;
            if (h_slot(joint.a) == h_slot(joint.b)) throw OraclePanic("joint: get2_mut identical indices");
            RigidBody* body_a = rbd_set.get(joint.a);
            RigidBody* body_b = rbd_set.get(joint.b);
            if (!body_a || !body_b) throw OraclePanic("joint: unwrap on None");

            Vec2 world_anchor_a = body_a->position + joint.anchor_a;
            Vec2 world_anchor_b = body_b->position + joint.anchor_b;
            Vec2 delta_position = world_anchor_b - world_anchor_a;
            float distance = length(delta_position);
            if (distance < 1e-6f) continue;

            float off_by = distance - joint.distance;
            Vec2 correction = off_by * delta_position / distance;

            if (!(body_a->calculated_mass > 0.0f) || !(body_b->calculated_mass > 0.0f))
                throw OraclePanic("assertion failed: calculated_mass > 0.0");
            float inv_mass_sum = (1.0f / body_a->calculated_mass) + (1.0f / body_b->calculated_mass);

            if (body_a->is_static()) {
                body_b->position -= inv_mass_sum * correction;
            } else if (body_b->is_static()) {
                body_a->position += inv_mass_sum * correction;
            } else {
                float ratio = (1.0f / body_a->calculated_mass) / inv_mass_sum;
                body_a->position += ratio * correction;
                body_b->position -= (1.0f - ratio) * correction;
            }

            float angle_a = std::atan2(delta_position.y, delta_position.x);
            float angle_b = -std::atan2(delta_position.y, -delta_position.x);
            float angle_diff = angle_b - angle_a - joint.target_angle;
            float rotation_correction = angle_diff * 0.5f;
            body_a->rotation += rotation_correction * dt;
            body_b->rotation -= rotation_correction * dt;
            if (std::isnan(body_a->rotation) || std::isnan(body_b->rotation) || std::isinf(body_a->rotation) ||
                std::isinf(body_b->rotation))
                throw OraclePanic("assertion failed: rotation finite");
        }
    }
}

// physics.rs:323-367
void Physics::update_objects(float dt) {
    for (uint32_t s = 0; s < rbd_set.slots(); ++s) {
        if (!rbd_set.alive(s)) continue;
        RigidBody& body = rbd_set.storage[s].value;
        if (body.is_static()) {
            body.position_old = body.position;
            body.acceleration = v2(0.f, 0.f);
            body.calculated_velocity = v2(0.f, 0.f);
            continue;
        }
        if (body.has_velocity_request) {  // velocity_request.take()
            body.has_velocity_request = false;
            body.position_old = body.position - body.velocity_request * dt;
        }
        Vec2 displacement = (body.position - body.position_old) * (dt / old_dt);
        old_dt = dt;  // inside the loop (Q2)
        if (maintain_spatial_hash) spatial_hash.move_point(rbd_set.handle_at(s), displacement);

        body.position_old = body.position;
        body.position += displacement + body.acceleration * dt * dt;

        body.angular_velocity += body.torque / body.inertia * dt;
        body.rotation += body.angular_velocity * dt;
        body.torque = 0.0f;
        body.acceleration = v2(0.f, 0.f);
        body.calculated_velocity = displacement / dt;
    }
    for (uint32_t s = 0; s < rbd_set.slots(); ++s) {
        if (!rbd_set.alive(s)) continue;
        RigidBody& body = rbd_set.storage[s].value;
        Affine2 tr = affine_from_angle_translation(body.rotation, body.position);  // rigid_body.rs:88-90
        for (Handle ch : body.colliders)
            if (Collider* c = col_set.get(ch)) c->absolute_transform = mul(tr, c->offset);
    }
}

// physics.rs:377-395
void Physics::apply_constraints() {
    for (const Constraint& constraint : constraints) {
        for (uint32_t s = 0; s < rbd_set.slots(); ++s) {
            if (!rbd_set.alive(s)) continue;
            RigidBody& body = rbd_set.storage[s].value;
            Vec2 obj = constraint.position;
            float radius = constraint.radius;
            Vec2 to_obj = body.position - obj;
            float dist = length(to_obj);
            float diff = radius;
            if (dist > diff) {
                Vec2 n = to_obj / dist;
                body.position = obj + n * diff;
            }
        }
    }
}

// physics.rs:397-422
void Physics::integrate(uint32_t nsub, float delta) {
    float step_delta = delta / (float)nsub;
    for (uint32_t i = 0; i < nsub; ++i) {
        apply_gravity();
        for (uint32_t s = 0; s < springs.slots(); ++s)
            if (springs.alive(s)) apply_spring(springs.storage[s].value);
        if (collisions_enabled) {
            if (use_spatial_hash) throw OraclePanic("spatial collisions not supported right now");  // :412
            if (use_grid_pairs)
                grid_collisions();
            else
                brute_force_collisions();
        }
        solve_fixed_joints(step_delta);
        update_objects(step_delta);
        apply_constraints();
        pair_substep_end.push_back(pair_a.size());
    }
}

}  // namespace oracle
