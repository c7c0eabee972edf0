// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// CPU oracle: a deliberately naive, sequential, AoS restatement of the reference's
// `blobs::Physics::step` hot path (darthdeus/blobs, `blobs/src/physics.rs` etc.).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this. The product library (libblobs_b200.so) never links or calls it.
//
// PARITY STATUS: "parity unpinned" for Physics::step — the reference ships no test or
// golden vector that calls Physics::step and the reference cannot
// be compiled in this image (no cargo/rustc). What IS pinned by the reference's own
// tests: SpatialHash (blobs/src/tests.rs:26-90) and the Affine2 composition
// (blobs/src/collider.rs:340-389); both are reproduced in tests/test_oracle_golden.py.
//
// Third-party arithmetic restated from the published crates (not under /root/reference):
//   glam 0.24.2   (Cargo.lock:603): Vec2 ops are per-component scalar IEEE f32, no FMA;
//                 dot = x*x' + y*y'; length = sqrt(dot); Vec2/f32 is a true divide;
//                 Mat2::from_angle = cols [cos, sin], [-sin, cos];
//                 Affine2*Affine2 = {M1*M2, M1*t2 + t1}; M*v = x_axis*v.x + y_axis*v.y.
//   thunderdome 0.6.1 (Cargo.lock:1694): generational arena, slot-ordered iteration,
//                 LIFO free list, generation bumped on reuse, to_bits = gen<<32 | slot.
// Build with: g++ -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile).
#pragma once
#include <cstdint>
#include <cmath>
#include <vector>
#include <optional>
#include <unordered_map>
#include <unordered_set>
#include <stdexcept>
#include <string>

namespace oracle {

// ---------------------------------------------------------------- glam::Vec2 (scalar f32)
struct Vec2 {
    float x = 0.f, y = 0.f;
};
inline Vec2 v2(float x, float y) { return Vec2{x, y}; }
inline Vec2 operator+(Vec2 a, Vec2 b) { return {a.x + b.x, a.y + b.y}; }
inline Vec2 operator-(Vec2 a, Vec2 b) { return {a.x - b.x, a.y - b.y}; }
inline Vec2 operator-(Vec2 a) { return {-a.x, -a.y}; }
inline Vec2 operator*(Vec2 a, float s) { return {a.x * s, a.y * s}; }
inline Vec2 operator*(float s, Vec2 a) { return {s * a.x, s * a.y}; }
inline Vec2 operator*(Vec2 a, Vec2 b) { return {a.x * b.x, a.y * b.y}; }
inline Vec2 operator/(Vec2 a, float s) { return {a.x / s, a.y / s}; }
inline Vec2 operator-(float s, Vec2 a) { return {s - a.x, s - a.y}; }  // `f32 - Vec2` (springs.rs:41)
inline Vec2& operator+=(Vec2& a, Vec2 b) { a = a + b; return a; }
inline Vec2& operator-=(Vec2& a, Vec2 b) { a = a - b; return a; }
inline float dot(Vec2 a, Vec2 b) { return (a.x * b.x) + (a.y * b.y); }
inline float length(Vec2 a) { return std::sqrt(dot(a, a)); }
inline float length_squared(Vec2 a) { return dot(a, a); }
inline float perp_dot(Vec2 a, Vec2 b) { return (a.x * b.y) - (a.y * b.x); }
inline bool is_nan(Vec2 a) { return std::isnan(a.x) || std::isnan(a.y); }

// ---------------------------------------------------------------- glam::Affine2
struct Mat2 {
    Vec2 x_axis{1.f, 0.f}, y_axis{0.f, 1.f};
};
inline Vec2 mul(const Mat2& m, Vec2 v) { return m.x_axis * v.x + m.y_axis * v.y; }
inline Mat2 mul(const Mat2& a, const Mat2& b) { return Mat2{mul(a, b.x_axis), mul(a, b.y_axis)}; }
inline Mat2 mat2_from_angle(float angle) {
    float s = std::sin(angle), c = std::cos(angle);  // f32::sin_cos -> libm sinf/cosf
    return Mat2{{c, s}, {-s, c}};
}
struct Affine2 {
    Mat2 matrix2;
    Vec2 translation;
};
inline Affine2 affine_from_angle_translation(float angle, Vec2 t) { return {mat2_from_angle(angle), t}; }
inline Affine2 affine_from_translation(Vec2 t) { return {Mat2{}, t}; }
inline Affine2 mul(const Affine2& a, const Affine2& b) {
    return Affine2{mul(a.matrix2, b.matrix2), mul(a.matrix2, b.translation) + a.translation};
}

// ---------------------------------------------------------------- thunderdome::Arena
// Index::to_bits = generation<<32 | slot; generation starts at 1 (NonZeroU32).
using Handle = uint64_t;
constexpr Handle NO_HANDLE = 0;  // generation 0 never exists
inline uint32_t h_slot(Handle h) { return (uint32_t)(h & 0xffffffffu); }
inline uint32_t h_gen(Handle h) { return (uint32_t)(h >> 32); }
inline Handle mk_handle(uint32_t slot, uint32_t gen) { return ((uint64_t)gen << 32) | slot; }

template <class T>
struct Arena {
    struct Entry {
        bool occupied = false;
        uint32_t generation = 0;
        int64_t next_free = -1;
        T value{};
    };
    std::vector<Entry> storage;
    int64_t first_free = -1;
    size_t len = 0;

    Handle insert(T v) {
        len++;
        if (first_free >= 0) {
            uint32_t slot = (uint32_t)first_free;
            Entry& e = storage[slot];
            first_free = e.next_free;
            uint32_t g = e.generation + 1;  // Generation::next (wraps to 1)
            if (g == 0) g = 1;
            e.occupied = true;
            e.generation = g;
            e.value = std::move(v);
            return mk_handle(slot, g);
        }
        Entry e;
        e.occupied = true;
        e.generation = 1;
        e.value = std::move(v);
        storage.push_back(std::move(e));
        return mk_handle((uint32_t)storage.size() - 1, 1);
    }
    T* get(Handle h) {
        uint32_t s = h_slot(h);
        if (s >= storage.size()) return nullptr;
        Entry& e = storage[s];
        if (!e.occupied || e.generation != h_gen(h)) return nullptr;
        return &e.value;
    }
    const T* get(Handle h) const { return const_cast<Arena*>(this)->get(h); }
    bool remove(Handle h) {
        if (!get(h)) return false;
        remove_slot(h_slot(h));
        return true;
    }
    void remove_slot(uint32_t s) {
        Entry& e = storage[s];
        e.occupied = false;
        e.value = T{};
        e.next_free = first_free;
        first_free = s;
        len--;
    }
    // Arena::clear == drain(): removes occupied slots in ascending order, each pushed on the
    // LIFO free list (restated from the published crate; not verifiable offline).
    void clear() {
        for (uint32_t s = 0; s < storage.size() && len > 0; ++s)
            if (storage[s].occupied) remove_slot(s);
    }
    size_t slots() const { return storage.size(); }
    bool alive(uint32_t s) const { return s < storage.size() && storage[s].occupied; }
    Handle handle_at(uint32_t s) const { return mk_handle(s, storage[s].generation); }
};

// ---------------------------------------------------------------- data model
enum BodyType : uint32_t { Dynamic = 0, Static = 1, KinematicPositionBased = 2, KinematicVelocityBased = 3 };

struct RigidBody {  // rigid_body.rs:41-74
    Vec2 position, position_old;
    Vec2 center_of_mass;
    float calculated_mass = 1.f;
    float gravity_mod = 1.f;
    float rotation = 0.f, angular_velocity = 0.f, torque = 0.f, inertia = 1.f;
    Vec2 scale{1.f, 1.f};
    Vec2 acceleration;
    bool has_velocity_request = false;
    Vec2 velocity_request;
    Vec2 calculated_velocity;
    std::vector<Handle> colliders;
    std::vector<Handle> connected_joints;
    uint64_t user_data_lo = 0, user_data_hi = 0;
    BodyType body_type = Dynamic;
    bool is_static() const { return body_type == Static; }
};

struct Collider {  // collider.rs:3-20
    Affine2 offset;
    Affine2 absolute_transform;
    uint64_t user_data_lo = 0, user_data_hi = 0;
    Handle parent = NO_HANDLE;  // Option<RigidBodyHandle>
    float radius = 0.5f;
    bool has_mass_override = false;
    float mass_override = 0.f;
    bool is_sensor = false;
    uint32_t memberships = 0xffffffffu, filter = 0xffffffffu;  // groups.rs:7-12
    float mass() const { return has_mass_override ? mass_override : radius * 2.0f; }  // collider.rs:40-42
    float inertia() const {                                                           // collider.rs:44-50
        float m = mass();
        float d = length(offset.translation);
        float in = 0.5f * m * (radius * radius);
        return in + m * (d * d);
    }
};

struct Spring {  // springs.rs:16-22
    Handle a = NO_HANDLE, b = NO_HANDLE;
    float rest_length = 0.f, stiffness = 0.f, damping = 0.f;
};

struct FixedJoint {  // joints.rs:6-19
    Handle a = NO_HANDLE, b = NO_HANDLE;
    Vec2 anchor_a, anchor_b;
    float distance = 0.f, target_angle = 0.f;
};

struct Constraint {  // lib.rs:189-193
    Vec2 position;
    float radius = 0.f;
};

struct CollisionEvent {  // lib.rs:146-153
    Handle col_a, col_b;
    Vec2 impact_vel_a, impact_vel_b;
};

// ---------------------------------------------------------------- SpatialHash (spatial.rs)
struct CellPoint {
    uint64_t id;
    Vec2 position;
    float radius;
};
struct SpatialHash {
    float cell_size;
    uint64_t next_id = 0;
    std::unordered_map<uint64_t, CellPoint> points;
    std::unordered_map<uint64_t, std::unordered_set<uint64_t>> grid;  // key = pack(cx, cy)
    std::vector<CellPoint> query_results;
    explicit SpatialHash(float cs = 2.0f) : cell_size(cs) {}
    static int32_t f2i_sat(float f) {  // Rust `as i32`: saturating, NaN -> 0
        if (std::isnan(f)) return 0;
        if (f >= 2147483648.0f) return INT32_MAX;
        if (f <= -2147483648.0f) return INT32_MIN;
        return (int32_t)f;
    }
    static uint64_t pack(int32_t cx, int32_t cy) { return ((uint64_t)(uint32_t)cx << 32) | (uint32_t)cy; }
    void get_cell_coords(Vec2 p, int32_t& cx, int32_t& cy) const {  // spatial.rs:57-62
        cx = f2i_sat(std::floor(p.x / cell_size));
        cy = f2i_sat(std::floor(p.y / cell_size));
    }
    uint64_t cell_key(Vec2 p) const {
        int32_t cx, cy;
        get_cell_coords(p, cx, cy);
        return pack(cx, cy);
    }
    uint64_t insert(Vec2 p, float r) {  // spatial.rs:64-69
        uint64_t id = next_id++;
        insert_with_id(id, p, r);
        return id;
    }
    void insert_with_id(uint64_t id, Vec2 p, float r) {  // spatial.rs:71-83
        grid[cell_key(p)].insert(id);
        points[id] = CellPoint{id, p, r};
    }
    bool remove(uint64_t id) {  // spatial.rs:85-100
        auto it = points.find(id);
        if (it == points.end()) return false;
        auto g = grid.find(cell_key(it->second.position));
        if (g != grid.end()) g->second.erase(id);
        points.erase(it);
        return true;
    }
    bool move_point(uint64_t id, Vec2 offset) {  // spatial.rs:102-137
        auto it = points.find(id);
        if (it == points.end()) return false;
        Vec2 oldp = it->second.position, newp = oldp + offset;
        uint64_t ok = cell_key(oldp), nk = cell_key(newp);
        if (ok != nk) {
            auto g = grid.find(ok);
            if (g != grid.end()) g->second.erase(id);
            grid[nk].insert(id);
        }
        it->second.position = newp;
        return true;
    }
    const std::vector<CellPoint>& query(Vec2 p, float qr) {  // spatial.rs:155-195
        int32_t x, y;
        get_cell_coords(p, x, y);
        query_results.clear();
        for (int dx = -1; dx <= 1; ++dx)
            for (int dy = -1; dy <= 1; ++dy) {
                auto g = grid.find(pack((int32_t)((uint32_t)x + (uint32_t)dx), (int32_t)((uint32_t)y + (uint32_t)dy)));
                if (g == grid.end()) continue;
                for (uint64_t pid : g->second) {
                    const CellPoint& pt = points.at(pid);
                    float d = qr + pt.radius;
                    if (length_squared(pt.position - p) <= d * d) query_results.push_back(pt);
                }
            }
        return query_results;
    }
};

struct OraclePanic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---------------------------------------------------------------- Physics (physics.rs)
struct Physics {
    Vec2 gravity;
    uint32_t substeps = 8;          // physics.rs:46
    uint32_t joint_iterations = 4;  // physics.rs:47
    Arena<RigidBody> rbd_set;
    Arena<Collider> col_set;
    Arena<FixedJoint> joints;
    Arena<Spring> springs;
    std::vector<Constraint> constraints;
    SpatialHash spatial_hash{2.0f};  // physics.rs:66
    bool use_spatial_hash = false;
    bool collisions_enabled = true;
    double accumulator = 0.0, time = 0.0;
    float old_dt = 1.0f;  // physics.rs:67

    // oracle instrumentation (not in the reference)
    bool maintain_spatial_hash = true;  // physics.rs:125-126,341 — costly, can be disabled for big N
    bool record_events = true;
    bool use_grid_pairs = false;  // "grid oracle": same arithmetic, cell-list pair enumeration
    std::vector<CollisionEvent> events;  // collision_send (physics.rs:304-311)
    std::vector<uint32_t> pair_a, pair_b;  // per substep contact pairs (slots), a > b
    std::vector<uint64_t> pair_substep_end;  // running pair count at the end of every substep
    uint64_t collisions_total = 0;  // perf_counter_inc("collisions", count) (physics.rs:316)
    uint64_t coincident_total = 0;

    explicit Physics(Vec2 g = {0.f, 0.f}, bool use_hash = false) : gravity(g), use_spatial_hash(use_hash) {}

    void reset();                       // physics.rs:71-76
    void step(double delta);            // physics.rs:78-82
    int fixed_step(double frame_time);  // physics.rs:84-99 (returns #integrate calls)
    Handle insert_rbd(const RigidBody& rbd);                              // physics.rs:121-128
    Handle insert_collider_with_parent(Collider col, Handle rbd_handle);  // physics.rs:130-149
    void remove_col(Handle h);                                            // physics.rs:159-161
    void remove_rbd(Handle h);                                            // physics.rs:163-172
    Handle create_fixed_joint(Handle a, Handle b, Vec2 anchor_a, Vec2 anchor_b);  // physics.rs:184-207
    Handle create_fixed_joint_with_distance(Handle a, Handle b, Vec2 anchor_a, Vec2 anchor_b, float distance);
    void update_mass_and_inertia(RigidBody& body);  // rigid_body.rs:96-128

    void integrate(uint32_t substeps, float delta);  // physics.rs:397-422
    void apply_gravity();                            // physics.rs:369-375
    void apply_spring(const Spring& s);              // springs.rs:25-47
    void brute_force_collisions();                   // physics.rs:241-317
    void grid_collisions();                          // same arithmetic, cell-list enumeration
    void solve_fixed_joints(float dt);               // physics.rs:424-477
    void update_objects(float dt);                   // physics.rs:323-367
    void apply_constraints();                        // physics.rs:377-395

   private:
    // returns false when the pair was skipped before the distance test
    void resolve_pair(uint32_t slot_a, uint32_t slot_b, uint64_t& count);
};

}  // namespace oracle
