//! Drop-in `blobs::Physics` over libblobs_b200.so. Same names / argument meaning / panics as the reference
//! (blobs/src/physics.rs); the world lives on the GPU, the shim keeps only handles and a per-frame host mirror.
//! NOT compiled in this repository's image (no cargo/rustc) — mechanical by construction, see INTEGRATION.md.
mod ffi;
use ffi::*;
use glam::{Affine2, Mat2, Vec2};
use std::ffi::CStr;
use std::sync::mpsc::{channel, Receiver, Sender};
use thunderdome::Index;

#[derive(Copy, Clone, Debug, Hash, PartialEq, Eq)] pub struct RigidBodyHandle(pub Index);
#[derive(Copy, Clone, Debug, Hash, PartialEq, Eq)] pub struct ColliderHandle(pub Index);
#[derive(Copy, Clone, Debug, Hash, PartialEq, Eq)] pub struct JointHandle(pub Index);
#[derive(Copy, Clone, Debug)] pub struct SpringHandle(pub Index);

#[derive(Copy, Clone, Debug)]
pub struct CollisionEvent { pub col_handle_a: ColliderHandle, pub col_handle_b: ColliderHandle, pub impact_vel_a: Vec2, pub impact_vel_b: Vec2 }
pub struct Constraint { pub position: Vec2, pub radius: f32 }
pub struct Spring { pub rigid_body_a: RigidBodyHandle, pub rigid_body_b: RigidBodyHandle, pub rest_length: f32, pub stiffness: f32, pub damping: f32 }

fn v(a: Vec2) -> BlobsVec2 { BlobsVec2 { x: a.x, y: a.y } }
fn g(a: BlobsVec2) -> Vec2 { Vec2::new(a.x, a.y) }
fn aff(a: Affine2) -> BlobsAffine2 { BlobsAffine2 { x_axis: v(a.matrix2.x_axis), y_axis: v(a.matrix2.y_axis), translation: v(a.translation) } }
fn idx(h: u64) -> Index { Index::from_bits(h).expect("valid handle bits") }

pub struct Physics {
    w: *mut BlobsWorld,
    /// The reference's `pub` tuning fields (physics.rs:6-27) stay plain fields, so `&mut physics.substeps` (the egui DragValue of
    /// demo/src/main.rs:263-273) compiles unchanged; `step` / `fixed_step` push them to the library before stepping.
    pub substeps: u32,
    pub joint_iterations: u32,
    pub gravity: Vec2,
    pub collisions_enabled: bool,
    /// `physics.springs.insert(Spring { .. })` / `.remove(index)` (demo/src/demos/joints.rs:59-66,79) on the library's spring arena
    pub springs: SpringArena,
    pub collision_send: Sender<CollisionEvent>,
    pub collision_recv: Receiver<CollisionEvent>,
}

/// Stand-in for `Arena<Spring>` (physics.rs:13): same `insert` / `remove` surface, same thunderdome indices.
pub struct SpringArena { w: *mut BlobsWorld }
impl SpringArena {
    pub fn insert(&mut self, s: Spring) -> Index {
        let mut h = 0u64;
        let rc = unsafe { blobs_spring_insert(self.w, s.rigid_body_a.0.to_bits(), s.rigid_body_b.0.to_bits(), s.rest_length, s.stiffness, s.damping, &mut h) };
        assert!(rc == BLOBS_OK, "{}", unsafe { CStr::from_ptr(blobs_last_error(self.w)) }.to_string_lossy());
        idx(h)
    }
    pub fn remove(&mut self, i: Index) -> Option<()> { (unsafe { blobs_spring_remove(self.w, i.to_bits()) } == BLOBS_OK).then_some(()) }
}

/// rigid_body.rs:19-26
#[derive(Copy, Clone, Debug, Default)]
pub struct RigidBodyData { pub position: Vec2, pub velocity: Vec2, pub angular_velocity: f32, pub center_of_mass: Vec2, pub mass: f32, pub rotation: f32 }

impl Physics {
    /// Physics::new (physics.rs:37-69)
    pub fn new(gravity: Vec2, use_spatial_hash: bool) -> Self {
        let p = BlobsParams { gravity: v(gravity), use_spatial_hash: use_spatial_hash as i32, device: -1, body_capacity_hint: 0, collider_capacity_hint: 0 };
        let mut w = std::ptr::null_mut();
        let rc = unsafe { blobs_world_create(&p, &mut w) };
        assert!(rc == BLOBS_OK, "blobs_world_create failed: {}", unsafe { CStr::from_ptr(blobs_last_error(std::ptr::null())) }.to_string_lossy());
        unsafe { blobs_record_contacts(w, BLOBS_RECORD_EVENTS, 1 << 20) }; // the reference always feeds collision_send
        let (collision_send, collision_recv) = channel();
        Self { w, substeps: 8, joint_iterations: 4, gravity, collisions_enabled: true, springs: SpringArena { w }, collision_send, collision_recv }   // physics.rs:46-47
    }
    /// the `pub` fields above -> library parameters (called by step / fixed_step)
    fn sync_params(&mut self) {
        unsafe {
            blobs_world_set_param(self.w, BLOBS_PARAM_SUBSTEPS, self.substeps as f64);
            blobs_world_set_param(self.w, BLOBS_PARAM_JOINT_ITERATIONS, self.joint_iterations as f64);
            blobs_world_set_param(self.w, BLOBS_PARAM_GRAVITY_X, self.gravity.x as f64);
            blobs_world_set_param(self.w, BLOBS_PARAM_GRAVITY_Y, self.gravity.y as f64);
            blobs_world_set_param(self.w, BLOBS_PARAM_COLLISIONS_ENABLED, self.collisions_enabled as i32 as f64);
        }
    }
    fn ck(&self, rc: i32) { if rc != BLOBS_OK { panic!("{}", unsafe { CStr::from_ptr(blobs_last_error(self.w)) }.to_string_lossy()); } }
    fn param(&self, id: i32) -> f64 { let mut x = 0.0; unsafe { blobs_world_get_param(self.w, id, &mut x) }; x }

    // `time` / `accumulator` (physics.rs:30-31) are advanced by the library: read-only here
    pub fn time(&self) -> f64 { self.param(BLOBS_PARAM_TIME) }
    pub fn push_constraint(&mut self, c: Constraint) { self.ck(unsafe { blobs_constraint_push(self.w, v(c.position), c.radius) }); }

    /// Physics::reset (physics.rs:71-76)
    pub fn reset(&mut self) { self.ck(unsafe { blobs_world_reset(self.w) }); }

    fn pump_events(&mut self) {
        let mut buf = vec![BlobsCollisionEvent::default(); 1 << 16];
        loop {
            let mut n = 0usize;
            self.ck(unsafe { blobs_events_drain(self.w, buf.as_mut_ptr(), buf.len(), &mut n) });
            for e in &buf[..n.min(buf.len())] {
                let _ = self.collision_send.send(CollisionEvent { col_handle_a: ColliderHandle(idx(e.col_handle_a)), col_handle_b: ColliderHandle(idx(e.col_handle_b)),
                                                                  impact_vel_a: g(e.impact_vel_a), impact_vel_b: g(e.impact_vel_b) });
            }
            if n <= buf.len() { break; }   // a partial drain leaves the rest queued (blobs_events_drain): keep pumping
        }
    }
    fn after_step(&mut self, st: &BlobsStepStats) {
        // the reference's channel is unbounded (physics.rs:22-23); the device-side recording is not: make an overflow loud
        assert!(st.events_dropped == 0, "blobs_b200: {} collision events did not fit the recording capacity (blobs_record_contacts)", st.events_dropped);
        self.pump_events();
    }
    /// Physics::step (physics.rs:78-82)
    pub fn step(&mut self, delta: f64) { let mut st = BlobsStepStats::default(); self.sync_params(); self.ck(unsafe { blobs_step(self.w, delta, &mut st) }); self.after_step(&st); }
    /// Physics::fixed_step (physics.rs:84-99)
    pub fn fixed_step(&mut self, frame_time: f64) { let mut st = BlobsStepStats::default(); self.sync_params(); self.ck(unsafe { blobs_fixed_step(self.w, frame_time, &mut st) }); self.after_step(&st); }

    /// insert_rbd (physics.rs:121-128); `RigidBody` is the builder output (rigid_body.rs:376-400)
    pub fn insert_rbd(&mut self, rbd: RigidBody) -> RigidBodyHandle {
        let d = BlobsBodyDesc { position: v(rbd.position), position_old: v(rbd.position_old), gravity_mod: rbd.gravity_mod, rotation: rbd.rotation, scale: v(rbd.scale),
            acceleration: v(rbd.acceleration), velocity_request: v(rbd.velocity_request.unwrap_or(Vec2::ZERO)), calculated_velocity: v(rbd.calculated_velocity),
            has_velocity_request: rbd.velocity_request.is_some() as i32, body_type: rbd.body_type as u32, user_data_lo: rbd.user_data as u64, user_data_hi: (rbd.user_data >> 64) as u64 };
        let mut h = 0u64;
        self.ck(unsafe { blobs_body_insert(self.w, &d, &mut h) });
        RigidBodyHandle(idx(h))
    }
    /// insert_collider_with_parent (physics.rs:130-149); panics "parent rigid body must exist when inserting collider"
    pub fn insert_collider_with_parent(&mut self, c: Collider, parent: RigidBodyHandle) -> ColliderHandle {
        let d = BlobsColliderDesc { offset: aff(c.offset), absolute_transform: aff(c.absolute_transform), radius: c.radius, mass_override: c.mass_override.unwrap_or(0.0),
            shape_radius: c.radius, has_mass_override: c.mass_override.is_some() as i32, is_sensor: c.flags.is_sensor as i32,
            memberships: c.collision_groups.memberships, filter: c.collision_groups.filter, user_data_lo: c.user_data as u64, user_data_hi: (c.user_data >> 64) as u64 };
        let mut h = 0u64;
        self.ck(unsafe { blobs_collider_insert(self.w, &d, parent.0.to_bits(), &mut h) });
        ColliderHandle(idx(h))
    }
    pub fn remove_rbd(&mut self, h: RigidBodyHandle) { unsafe { blobs_body_remove(self.w, h.0.to_bits()) }; }     // physics.rs:163-172 (missing body: event only)
    pub fn remove_col(&mut self, h: ColliderHandle) { unsafe { blobs_collider_remove(self.w, h.0.to_bits()) }; } // physics.rs:159-161
    /// What SpatialHash::query (spatial.rs:155-195) and QueryPipeline::intersection_with_shape (lib.rs:177-187, a stub in the
    /// reference) are for: every collider whose snapshot circle touches the query circle and passes `filter`, ascending slot order.
    pub fn query_circle(&mut self, position: Vec2, radius: f32, filter: QueryFilter) -> Vec<ColliderHandle> {
        let f = BlobsQueryFilter { flags: filter.flags, has_groups: filter.groups.is_some() as i32,
            memberships: filter.groups.map(|g| g.memberships).unwrap_or(0), filter: filter.groups.map(|g| g.filter).unwrap_or(0),
            exclude_collider: filter.exclude_collider.map(|h| h.0.to_bits()).unwrap_or(0),
            exclude_rigid_body: filter.exclude_rigid_body.map(|h| h.0.to_bits()).unwrap_or(0), batch_world: 0, reserved: 0 };
        let (c, r) = ([position.x, position.y], [radius]);
        let mut off = [0u64; 2];
        let mut n = 0usize;
        let mut hits = vec![0u64; 64];
        loop {
            let rc = unsafe { blobs_query_circles(self.w, 1, c.as_ptr(), r.as_ptr(), &f, off.as_mut_ptr(), hits.as_mut_ptr(), hits.len(), &mut n) };
            if rc == BLOBS_ERR_CAPACITY && n > hits.len() { hits.resize(n, 0); continue; }
            self.ck(rc);
            break;
        }
        hits[..n].iter().map(|&h| ColliderHandle(idx(h))).collect()
    }
    /// Physics::debug_data (physics.rs:479-481): one library call, lists in arena order (debug.rs:34-91)
    pub fn debug_data(&mut self) -> DebugData {
        let mut c = BlobsDebugCounts::default();
        self.ck(unsafe { blobs_debug_counts(self.w, &mut c) });
        let (mut bx, mut jx) = (vec![0f32; 6 * c.bodies as usize], vec![0f32; 4 * c.joints as usize]);
        let (mut cx, mut cr, mut sx) = (vec![0f32; 6 * c.colliders as usize], vec![0f32; c.colliders as usize], vec![0f32; 4 * c.springs as usize]);
        self.ck(unsafe { blobs_debug_data(self.w, bx.as_mut_ptr(), jx.as_mut_ptr(), cx.as_mut_ptr(), cr.as_mut_ptr(), sx.as_mut_ptr(), &c) });
        let aff = |r: &[f32]| Affine2::from_mat2_translation(Mat2::from_cols(Vec2::new(r[0], r[1]), Vec2::new(r[2], r[3])), Vec2::new(r[4], r[5]));
        DebugData {
            bodies: bx.chunks(6).map(|r| DebugRigidBody { transform: aff(r) }).collect(),
            joints: jx.chunks(4).map(|r| DebugJoint { body_a: Vec2::new(r[0], r[1]), body_b: Vec2::new(r[2], r[3]) }).collect(),
            colliders: cx.chunks(6).zip(cr.iter()).map(|(r, &radius)| DebugCollider { transform: aff(r), radius }).collect(),
            springs: sx.chunks(4).map(|r| DebugSpring { body_a: Vec2::new(r[0], r[1]), body_b: Vec2::new(r[2], r[3]) }).collect(),
        }
    }
    pub fn rbd_count(&self) -> usize { let mut n = 0u64; unsafe { blobs_body_count(self.w, &mut n) }; n as usize }
    pub fn get_rbd_state(&mut self, h: RigidBodyHandle) -> Option<BlobsBodyState> { let mut s = BlobsBodyState::default(); (unsafe { blobs_body_get(self.w, h.0.to_bits(), &mut s) } == BLOBS_OK).then_some(s) }
    /// get_rbd (physics.rs:105-107): the reference hands out `&RigidBody`; the body lives in HBM, so this is a snapshot BY VALUE with the
    /// same `pub` fields and read-only methods (`physics.get_rbd(h).unwrap().position` reads the same either way)
    pub fn get_rbd(&mut self, h: RigidBodyHandle) -> Option<RigidBodyMirror> { self.get_rbd_state(h).map(|s| RigidBodyMirror::from_state(&s)) }
    /// get_rbd_data (physics.rs:101-103)
    pub fn get_rbd_data(&mut self, h: RigidBodyHandle) -> Option<RigidBodyData> {
        self.get_rbd_state(h).map(|s| RigidBodyData { position: g(s.position), velocity: g(s.calculated_velocity), angular_velocity: s.angular_velocity,
                                                      center_of_mass: g(s.center_of_mass), mass: s.calculated_mass, rotation: s.rotation })
    }
    /// get_col (physics.rs:117-119), by value like get_rbd
    pub fn get_col(&mut self, h: ColliderHandle) -> Option<Collider> {
        let mut s = BlobsColliderState::default();
        (unsafe { blobs_collider_get(self.w, h.0.to_bits(), &mut s) } == BLOBS_OK).then(|| {
            let a = |t: BlobsAffine2| Affine2 { matrix2: Mat2::from_cols(g(t.x_axis), g(t.y_axis)), translation: g(t.translation) };
            Collider { offset: a(s.desc.offset), absolute_transform: a(s.desc.absolute_transform), user_data: (s.desc.user_data_lo as u128) | ((s.desc.user_data_hi as u128) << 64),
                       radius: s.desc.radius, mass_override: (s.desc.has_mass_override != 0).then_some(s.desc.mass_override),
                       flags: ColliderFlags { is_sensor: s.desc.is_sensor != 0 },
                       collision_groups: InteractionGroups { memberships: s.desc.memberships, filter: s.desc.filter } }
        })
    }
    pub fn rbd_position(&mut self, h: RigidBodyHandle) -> Option<Vec2> { self.get_rbd_state(h).map(|s| g(s.position)) }              // physics.rs:151-153
    pub fn col_position(&mut self, h: ColliderHandle) -> Option<Vec2> { let mut s = BlobsColliderState::default(); (unsafe { blobs_collider_get(self.w, h.0.to_bits(), &mut s) } == BLOBS_OK).then(|| g(s.desc.absolute_transform.translation)) }
    pub fn update_rigid_body_position(&mut self, id: u64, offset: Vec2) { unsafe { blobs_body_translate(self.w, id, v(offset)) }; }   // physics.rs:174-182
    /// get_mut_rbd (physics.rs:109-111). The body lives in HBM, so the `&mut RigidBody` of the reference becomes a guard that
    /// derefs to a host mirror of the body; when the guard is dropped the fields that changed are written back with one
    /// blobs_body_set (field mask = dirty set). Every RigidBody method of the reference (rigid_body.rs:130-214) then runs
    /// unchanged on the mirror - same glam arithmetic, hence the same bits.
    pub fn get_mut_rbd(&mut self, h: RigidBodyHandle) -> Option<RigidBodyMut<'_>> {
        let before = self.get_rbd_state(h)?;
        Some(RigidBodyMut { physics: self, handle: h, before, body: RigidBodyMirror::from_state(&before) })
    }
    /// lower-level form of the same: write back the fields named in `mask`
    pub fn set_rbd_state(&mut self, h: RigidBodyHandle, s: &BlobsBodyState, mask: u32) { self.ck(unsafe { blobs_body_set(self.w, h.0.to_bits(), s, mask) }); }
    /// create_fixed_joint (physics.rs:184-207)
    pub fn create_fixed_joint(&mut self, a: RigidBodyHandle, b: RigidBodyHandle, anchor_a: Vec2, anchor_b: Vec2) -> JointHandle { self.create_fixed_joint_with_distance(a, b, anchor_a, anchor_b, f32::NAN) }
    /// create_fixed_joint_with_distance (physics.rs:209-239)
    pub fn create_fixed_joint_with_distance(&mut self, a: RigidBodyHandle, b: RigidBodyHandle, anchor_a: Vec2, anchor_b: Vec2, distance: f32) -> JointHandle {
        let mut h = 0u64; self.ck(unsafe { blobs_joint_insert(self.w, a.0.to_bits(), b.0.to_bits(), v(anchor_a), v(anchor_b), distance, &mut h) }); JointHandle(idx(h))
    }
    /// physics.springs.insert(Spring{..}) (demo/src/demos/joints.rs:59-66)
    pub fn insert_spring(&mut self, s: Spring) -> SpringHandle { let mut h = 0u64; self.ck(unsafe { blobs_spring_insert(self.w, s.rigid_body_a.0.to_bits(), s.rigid_body_b.0.to_bits(), s.rest_length, s.stiffness, s.damping, &mut h) }); SpringHandle(idx(h)) }
}
impl Drop for Physics { fn drop(&mut self) { unsafe { blobs_world_destroy(self.w) }; } }

// ---- builders: field-for-field the reference's (rigid_body.rs:287-401, collider.rs:199-284) ------------------------
#[derive(Copy, Clone, Debug, PartialEq, Eq)] pub enum RigidBodyType { Dynamic = 0, Static = 1, KinematicPositionBased = 2, KinematicVelocityBased = 3 }
#[derive(Copy, Clone, Debug, PartialEq, Eq)] pub struct InteractionGroups { pub memberships: u32, pub filter: u32 }
impl Default for InteractionGroups { fn default() -> Self { Self { memberships: u32::MAX, filter: u32::MAX } } }
#[derive(Copy, Clone, Debug, Default, PartialEq, Eq)] pub struct ColliderFlags { pub is_sensor: bool }

pub struct RigidBody { pub position: Vec2, pub position_old: Vec2, pub gravity_mod: f32, pub rotation: f32, pub scale: Vec2, pub acceleration: Vec2,
    pub velocity_request: Option<Vec2>, pub calculated_velocity: Vec2, pub user_data: u128, pub body_type: RigidBodyType }
pub struct RigidBodyBuilder(RigidBody);
impl RigidBodyBuilder {
    pub fn new() -> Self { Self(RigidBody { position: Vec2::ZERO, position_old: Vec2::ZERO, gravity_mod: 1.0, rotation: 0.0, scale: Vec2::ONE, acceleration: Vec2::ZERO,
        velocity_request: None, calculated_velocity: Vec2::ZERO, user_data: 0, body_type: RigidBodyType::Dynamic }) }
    pub fn position(mut self, p: Vec2) -> Self { self.0.position_old = p; self.0.position = p; self }
    pub fn gravity_mod(mut self, x: f32) -> Self { self.0.gravity_mod = x; self }
    pub fn rotation(mut self, x: f32) -> Self { self.0.rotation = x; self }
    pub fn scale(mut self, x: Vec2) -> Self { self.0.scale = x; self }
    pub fn acceleration(mut self, x: Vec2) -> Self { self.0.acceleration = x; self }
    pub fn velocity_request(mut self, x: Vec2) -> Self { self.0.velocity_request = Some(x); self }
    pub fn calculated_velocity(mut self, x: Vec2) -> Self { self.0.calculated_velocity = x; self }
    pub fn user_data(mut self, x: u128) -> Self { self.0.user_data = x; self }
    pub fn body_type(mut self, x: RigidBodyType) -> Self { self.0.body_type = x; self }
    pub fn build(self) -> RigidBody { self.0 }
}
pub struct Collider { pub offset: Affine2, pub absolute_transform: Affine2, pub user_data: u128, pub radius: f32, pub mass_override: Option<f32>,
    pub flags: ColliderFlags, pub collision_groups: InteractionGroups }
pub struct ColliderBuilder(Collider);
impl ColliderBuilder {
    pub fn new() -> Self { Self(Collider { offset: Affine2::IDENTITY, absolute_transform: Affine2::IDENTITY, user_data: 0, radius: 0.5, mass_override: None,
        flags: ColliderFlags::default(), collision_groups: InteractionGroups::default() }) }
    pub fn offset(mut self, x: Affine2) -> Self { self.0.offset = x; self }
    pub fn absolute_transform(mut self, x: Affine2) -> Self { self.0.absolute_transform = x; self }
    pub fn mass_override(mut self, x: f32) -> Self { self.0.mass_override = Some(x); self }
    pub fn user_data(mut self, x: u128) -> Self { self.0.user_data = x; self }
    pub fn radius(mut self, x: f32) -> Self { self.0.radius = x; self }
    pub fn flags(mut self, x: ColliderFlags) -> Self { self.0.flags = x; self }
    pub fn collision_groups(mut self, x: InteractionGroups) -> Self { self.0.collision_groups = x; self }
    pub fn build(self) -> Collider { self.0 }
}
#[allow(dead_code)] fn _unused(_: Mat2) {}

/// Host mirror of one RigidBody (rigid_body.rs:41-74): the `pub` fields plus the reference's methods, verbatim semantics.
#[derive(Copy, Clone, Debug)]
pub struct RigidBodyMirror {
    pub position: Vec2, pub position_old: Vec2, pub center_of_mass: Vec2, pub scale: Vec2, pub acceleration: Vec2,
    pub velocity_request: Option<Vec2>, pub calculated_velocity: Vec2, pub calculated_mass: f32, pub gravity_mod: f32,
    pub rotation: f32, pub angular_velocity: f32, pub torque: f32, pub inertia: f32, pub user_data: u128, pub body_type: RigidBodyType,
}
impl RigidBodyMirror {
    fn from_state(s: &BlobsBodyState) -> Self {
        let body_type = match s.body_type { 1 => RigidBodyType::Static, 2 => RigidBodyType::KinematicPositionBased, 3 => RigidBodyType::KinematicVelocityBased, _ => RigidBodyType::Dynamic };
        Self { position: g(s.position), position_old: g(s.position_old), center_of_mass: g(s.center_of_mass), scale: g(s.scale), acceleration: g(s.acceleration),
               velocity_request: (s.has_velocity_request != 0).then(|| g(s.velocity_request)), calculated_velocity: g(s.calculated_velocity),
               calculated_mass: s.calculated_mass, gravity_mod: s.gravity_mod, rotation: s.rotation, angular_velocity: s.angular_velocity, torque: s.torque,
               inertia: s.inertia, user_data: (s.user_data_lo as u128) | ((s.user_data_hi as u128) << 64), body_type }
    }
    pub fn is_static(&self) -> bool { self.body_type == RigidBodyType::Static }                                  // rigid_body.rs:211-213
    pub fn is_dynamic(&self) -> bool { self.body_type == RigidBodyType::Dynamic }                                // :198-200
    pub fn is_kinematic(&self) -> bool { matches!(self.body_type, RigidBodyType::KinematicPositionBased | RigidBodyType::KinematicVelocityBased) } // :202-205
    pub fn get_velocity(&self) -> Vec2 { self.calculated_velocity }                                              // :186-188
    pub fn set_velocity(&mut self, velocity: Vec2) { self.velocity_request = Some(velocity); }                   // :182-184
    pub fn add_velocity(&mut self, velocity: Vec2) { self.set_velocity(self.get_velocity() + velocity); }        // :151-153
    pub fn apply_impulse(&mut self, impulse: Vec2) { if !self.is_static() { self.add_velocity(impulse / self.calculated_mass); } }   // :130-135
    pub fn apply_impulse_at_point(&mut self, impulse: Vec2, world_point: Vec2) {                                 // :137-149
        if !self.is_static() {
            self.apply_impulse(impulse);
            let lever_arm = world_point - (self.position + self.center_of_mass);
            self.angular_velocity += lever_arm.perp_dot(impulse) / self.inertia;
        }
    }
    pub fn apply_force(&mut self, force: Vec2) { if !self.is_static() { self.acceleration += force / self.calculated_mass; } }       // :155-160
    pub fn apply_force_at_point(&mut self, force: Vec2, world_point: Vec2) {                                     // :162-172
        if !self.is_static() {
            self.apply_force(force);
            let lever_arm = world_point - (self.position + self.center_of_mass);
            self.torque += lever_arm.perp_dot(force);
        }
    }
    pub fn apply_torque_at_point(&mut self, force: Vec2, world_point: Vec2) {                                    // :174-180
        if !self.is_static() { let lever_arm = world_point - (self.position + self.center_of_mass); self.torque += lever_arm.perp_dot(force); }
    }
    pub fn accelerate(&mut self, a: Vec2) { self.acceleration += a; }                                            // :207-209
}

/// The guard `get_mut_rbd` returns: `Deref`/`DerefMut` to the mirror, write-back of the dirty fields on drop.
pub struct RigidBodyMut<'a> { physics: &'a mut Physics, handle: RigidBodyHandle, before: BlobsBodyState, body: RigidBodyMirror }
impl std::ops::Deref for RigidBodyMut<'_> { type Target = RigidBodyMirror; fn deref(&self) -> &RigidBodyMirror { &self.body } }
impl std::ops::DerefMut for RigidBodyMut<'_> { fn deref_mut(&mut self) -> &mut RigidBodyMirror { &mut self.body } }
impl Drop for RigidBodyMut<'_> {
    fn drop(&mut self) {
        let (b, o) = (&self.body, RigidBodyMirror::from_state(&self.before));
        let bits = |a: Vec2, c: Vec2| a.x.to_bits() != c.x.to_bits() || a.y.to_bits() != c.y.to_bits();
        let mut s = self.before;
        let mut mask = 0u32;
        if bits(b.position, o.position) { s.position = v(b.position); mask |= BLOBS_BODY_POSITION; }
        if bits(b.position_old, o.position_old) { s.position_old = v(b.position_old); mask |= BLOBS_BODY_POSITION_OLD; }
        if bits(b.acceleration, o.acceleration) { s.acceleration = v(b.acceleration); mask |= BLOBS_BODY_ACCELERATION; }
        if bits(b.calculated_velocity, o.calculated_velocity) { s.calculated_velocity = v(b.calculated_velocity); mask |= BLOBS_BODY_CALC_VELOCITY; }
        if bits(b.scale, o.scale) { s.scale = v(b.scale); mask |= BLOBS_BODY_SCALE; }
        if bits(b.center_of_mass, o.center_of_mass) { s.center_of_mass = v(b.center_of_mass); mask |= BLOBS_BODY_CENTER_OF_MASS; }
        if b.velocity_request.map(|q| (q.x.to_bits(), q.y.to_bits())) != o.velocity_request.map(|q| (q.x.to_bits(), q.y.to_bits())) {
            s.has_velocity_request = b.velocity_request.is_some() as i32;
            s.velocity_request = v(b.velocity_request.unwrap_or(Vec2::ZERO));
            mask |= BLOBS_BODY_VELOCITY_REQUEST;
        }
        if b.rotation.to_bits() != o.rotation.to_bits() { s.rotation = b.rotation; mask |= BLOBS_BODY_ROTATION; }
        if b.angular_velocity.to_bits() != o.angular_velocity.to_bits() { s.angular_velocity = b.angular_velocity; mask |= BLOBS_BODY_ANGULAR_VELOCITY; }
        if b.torque.to_bits() != o.torque.to_bits() { s.torque = b.torque; mask |= BLOBS_BODY_TORQUE; }
        if b.calculated_mass.to_bits() != o.calculated_mass.to_bits() { s.calculated_mass = b.calculated_mass; mask |= BLOBS_BODY_MASS; }
        if b.inertia.to_bits() != o.inertia.to_bits() { s.inertia = b.inertia; mask |= BLOBS_BODY_INERTIA; }
        if b.gravity_mod.to_bits() != o.gravity_mod.to_bits() { s.gravity_mod = b.gravity_mod; mask |= BLOBS_BODY_GRAVITY_MOD; }
        if b.body_type != o.body_type { s.body_type = b.body_type as u32; mask |= BLOBS_BODY_TYPE; }
        if b.user_data != o.user_data { s.user_data_lo = b.user_data as u64; s.user_data_hi = (b.user_data >> 64) as u64; mask |= BLOBS_BODY_USER_DATA; }
        if mask != 0 { self.physics.set_rbd_state(self.handle, &s, mask); }
    }
}

/// perf_counters.rs:52-87 - same free functions, backed by the library's process-global registry (blobs_step feeds "collisions")
pub mod perf_counters {
    use super::ffi::*;
    use std::ffi::CString;
    pub fn perf_counter(counter_name: &str, count: u64) { let n = CString::new(counter_name).unwrap(); unsafe { blobs_perf_counter(n.as_ptr(), count) } }
    pub fn perf_counter_inc(counter_name: &str, inc: u64) { let n = CString::new(counter_name).unwrap(); unsafe { blobs_perf_counter_inc(n.as_ptr(), inc) } }
    pub fn perf_counters_new_frame(delta: f64) { unsafe { blobs_perf_counters_new_frame(delta) } }
    pub fn reset_perf_counters() { unsafe { blobs_perf_counters_reset() } }
    pub fn get_perf_counter(counter_name: &str) -> (u64, f64) {
        let n = CString::new(counter_name).unwrap();
        let (mut c, mut a) = (0u64, 0f64);
        unsafe { blobs_perf_counter_get(n.as_ptr(), &mut c, &mut a) };
        (c, a)
    }
    /// what the demo's perf panel iterates (demo/src/main.rs:291-300): (name, count, decayed_average)
    pub fn counters() -> Vec<(String, u64, f64)> {
        let mut out = Vec::new();
        for i in 0..unsafe { blobs_perf_counter_count() } {
            let mut name = [0 as std::os::raw::c_char; 256];
            let (mut c, mut a) = (0u64, 0f64);
            if unsafe { blobs_perf_counter_at(i, name.as_mut_ptr(), name.len(), &mut c, &mut a) } == BLOBS_OK {
                out.push((unsafe { std::ffi::CStr::from_ptr(name.as_ptr()) }.to_string_lossy().into_owned(), c, a));
            }
        }
        out
    }
}

/// debug.rs:6-32
pub struct DebugRigidBody { pub transform: Affine2 }
pub struct DebugCollider { pub transform: Affine2, pub radius: f32 }
pub struct DebugJoint { pub body_a: Vec2, pub body_b: Vec2 }
pub struct DebugSpring { pub body_a: Vec2, pub body_b: Vec2 }
pub struct DebugData { pub bodies: Vec<DebugRigidBody>, pub joints: Vec<DebugJoint>, pub colliders: Vec<DebugCollider>, pub springs: Vec<DebugSpring> }

/// query_filter.rs:74-87 (the closure predicate of the reference is applied by the caller on the returned handles);
/// flags: QueryFilterFlags bits (query_filter.rs:6-25), e.g. EXCLUDE_SENSORS = 1 << 4
#[derive(Copy, Clone, Default)]
pub struct QueryFilter { pub flags: u32, pub groups: Option<InteractionGroups>, pub exclude_collider: Option<ColliderHandle>, pub exclude_rigid_body: Option<RigidBodyHandle> }
