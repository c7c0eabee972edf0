"""Host-side mirror of the reference's public Rust API (blobs/src/physics.rs, rigid_body.rs, collider.rs, joints.rs,
springs.rs, groups.rs, lib.rs) over the C ABI — same names, argument meaning and error behaviour, so code and tests read
like the reference's. All state lives in the GPU world; this module only builds descriptors and forwards calls."""
import math
from dataclasses import dataclass, field

import numpy as np

from . import _abi as A
from .world import BlobsError, World


class RigidBodyType:  # rigid_body.rs:221-242
    Dynamic, Static, KinematicPositionBased, KinematicVelocityBased = range(4)


@dataclass
class InteractionGroups:  # groups.rs:7-57
    memberships: int = 0xFFFFFFFF
    filter: int = 0xFFFFFFFF

    @staticmethod
    def all():
        return InteractionGroups()

    @staticmethod
    def none():
        return InteractionGroups(0, 0)

    def test(self, rhs):
        return (self.memberships & rhs.filter) != 0 and (rhs.memberships & self.filter) != 0


def groups(memberships, filter):  # groups.rs:3-5
    return InteractionGroups(int(memberships), int(filter))


@dataclass
class Affine2:  # glam::Affine2 (columns)
    x_axis: tuple = (1.0, 0.0)
    y_axis: tuple = (0.0, 1.0)
    translation: tuple = (0.0, 0.0)

    @staticmethod
    def from_translation(t):
        return Affine2(translation=(float(t[0]), float(t[1])))

    @staticmethod
    def from_angle_translation(angle, t):
        s, c = math.sin(angle), math.cos(angle)
        return Affine2((c, s), (-s, c), (float(t[0]), float(t[1])))


@dataclass
class RigidBody:  # the builder's output (rigid_body.rs:376-400)
    position: tuple = (0.0, 0.0)
    position_old: tuple = (0.0, 0.0)
    gravity_mod: float = 1.0
    rotation: float = 0.0
    scale: tuple = (1.0, 1.0)
    acceleration: tuple = (0.0, 0.0)
    velocity_request: tuple = None
    calculated_velocity: tuple = (0.0, 0.0)
    user_data: int = 0
    body_type: int = RigidBodyType.Dynamic


class RigidBodyBuilder:  # rigid_body.rs:287-401
    def __init__(self):
        self._b = RigidBody()

    def position(self, p):
        self._b.position = self._b.position_old = (float(p[0]), float(p[1]))
        return self

    def gravity_mod(self, x):
        self._b.gravity_mod = float(x)
        return self

    def rotation(self, x):
        self._b.rotation = float(x)
        return self

    def scale(self, x):
        self._b.scale = tuple(x)
        return self

    def acceleration(self, x):
        self._b.acceleration = tuple(x)
        return self

    def velocity_request(self, x):
        self._b.velocity_request = tuple(x)
        return self

    def calculated_velocity(self, x):
        self._b.calculated_velocity = tuple(x)
        return self

    def user_data(self, x):
        self._b.user_data = int(x)
        return self

    def body_type(self, x):
        self._b.body_type = int(x)
        return self

    def build(self):
        return self._b


@dataclass
class ColliderFlags:
    is_sensor: bool = False


@dataclass
class Collider:  # collider.rs:3-20
    offset: Affine2 = field(default_factory=Affine2)
    absolute_transform: Affine2 = field(default_factory=Affine2)
    user_data: int = 0
    radius: float = 0.5
    mass_override: float = None
    flags: ColliderFlags = field(default_factory=ColliderFlags)
    collision_groups: InteractionGroups = field(default_factory=InteractionGroups)


class ColliderBuilder:  # collider.rs:199-284
    def __init__(self):
        self._c = Collider()

    def offset(self, x):
        self._c.offset = x
        return self

    def absolute_transform(self, x):
        self._c.absolute_transform = x
        return self

    def mass_override(self, x):
        self._c.mass_override = float(x)
        return self

    def user_data(self, x):
        self._c.user_data = int(x)
        return self

    def radius(self, x):
        self._c.radius = float(x)
        return self

    def flags(self, x):
        self._c.flags = x
        return self

    def collision_groups(self, x):
        self._c.collision_groups = x
        return self

    def build(self):
        return self._c


@dataclass
class Spring:  # springs.rs:16-22
    rigid_body_a: int
    rigid_body_b: int
    rest_length: float
    stiffness: float
    damping: float


@dataclass
class Constraint:  # lib.rs:189-193
    position: tuple
    radius: float


@dataclass
class CollisionEvent:  # lib.rs:146-153
    col_handle_a: int
    col_handle_b: int
    impact_vel_a: tuple
    impact_vel_b: tuple


_f32 = np.float32


class RigidBodyMut:
    """What `physics.get_mut_rbd(handle)` hands out (physics.rs:109-111): a host mirror of ONE RigidBody (rigid_body.rs:41-74)
    whose `pub` fields can be assigned and whose methods are the reference's (rigid_body.rs:130-214), evaluated on the host
    in f32 exactly as written there. The body itself lives in HBM: `commit()` (or leaving the `with` block) compares the
    mirror with what was downloaded and writes back only the fields that changed (one blobs_body_set with a field mask) -
    the dirty tracking a `&mut RigidBody` borrow needs when the storage is on the GPU."""

    _VEC = {"position": A.BODY_POSITION, "position_old": A.BODY_POSITION_OLD, "acceleration": A.BODY_ACCELERATION,
            "calculated_velocity": A.BODY_CALC_VELOCITY, "scale": A.BODY_SCALE, "center_of_mass": A.BODY_CENTER_OF_MASS}
    _SCALAR = {"rotation": A.BODY_ROTATION, "angular_velocity": A.BODY_ANGULAR_VELOCITY, "torque": A.BODY_TORQUE,
               "calculated_mass": A.BODY_MASS, "inertia": A.BODY_INERTIA, "gravity_mod": A.BODY_GRAVITY_MOD}

    def __init__(self, world, handle, state):
        self._w, self._handle = world, handle
        self._st0 = state.copy()
        for k in self._VEC:
            setattr(self, k, (_f32(state[k]["x"]), _f32(state[k]["y"])))
        for k in self._SCALAR:
            setattr(self, k, _f32(state[k]))
        self.velocity_request = (_f32(state["velocity_request"]["x"]), _f32(state["velocity_request"]["y"])) if state["has_velocity_request"] else None
        self.body_type = int(state["body_type"])
        self.user_data = int(state["user_data_lo"]) | (int(state["user_data_hi"]) << 64)

    # ---- rigid_body.rs:130-214, f32 arithmetic in the reference's order (glam Vec2 ops are per-component scalar ops)
    def is_static(self):
        return self.body_type == RigidBodyType.Static

    def is_dynamic(self):
        return self.body_type == RigidBodyType.Dynamic

    def is_kinematic(self):
        return self.body_type in (RigidBodyType.KinematicPositionBased, RigidBodyType.KinematicVelocityBased)

    def get_velocity(self):  # :186-188
        return self.calculated_velocity

    def set_velocity(self, velocity):  # :182-184
        self.velocity_request = (_f32(velocity[0]), _f32(velocity[1]))

    def add_velocity(self, velocity):  # :151-153
        v = self.get_velocity()
        self.set_velocity((v[0] + _f32(velocity[0]), v[1] + _f32(velocity[1])))

    def apply_impulse(self, impulse):  # :130-135
        if not self.is_static():
            self.add_velocity((_f32(impulse[0]) / self.calculated_mass, _f32(impulse[1]) / self.calculated_mass))

    def _lever_arm(self, world_point):
        cx, cy = self.position[0] + self.center_of_mass[0], self.position[1] + self.center_of_mass[1]
        return _f32(world_point[0]) - cx, _f32(world_point[1]) - cy

    @staticmethod
    def _perp_dot(a, b):  # glam Vec2::perp_dot: a.x * b.y - a.y * b.x
        return a[0] * _f32(b[1]) - a[1] * _f32(b[0])

    def apply_impulse_at_point(self, impulse, world_point):  # :137-149
        if not self.is_static():
            self.apply_impulse(impulse)
            self.angular_velocity = self.angular_velocity + self._perp_dot(self._lever_arm(world_point), impulse) / self.inertia

    def apply_force(self, force):  # :155-160
        if not self.is_static():
            self.acceleration = (self.acceleration[0] + _f32(force[0]) / self.calculated_mass,
                                 self.acceleration[1] + _f32(force[1]) / self.calculated_mass)

    def apply_force_at_point(self, force, world_point):  # :162-172
        if not self.is_static():
            self.apply_force(force)
            self.torque = self.torque + self._perp_dot(self._lever_arm(world_point), force)

    def apply_torque_at_point(self, force, world_point):  # :174-180
        if not self.is_static():
            self.torque = self.torque + self._perp_dot(self._lever_arm(world_point), force)

    def accelerate(self, a):  # :207-209
        self.acceleration = (self.acceleration[0] + _f32(a[0]), self.acceleration[1] + _f32(a[1]))

    # ---- write-back
    def commit(self):
        st, mask = self._st0.copy(), 0
        for k, bit in self._VEC.items():
            v = getattr(self, k)
            if _f32(v[0]).tobytes() != _f32(self._st0[k]["x"]).tobytes() or _f32(v[1]).tobytes() != _f32(self._st0[k]["y"]).tobytes():
                st[k]["x"], st[k]["y"] = v
                mask |= bit
        for k, bit in self._SCALAR.items():
            if _f32(getattr(self, k)).tobytes() != _f32(self._st0[k]).tobytes():
                st[k] = getattr(self, k)
                mask |= bit
        had = bool(self._st0["has_velocity_request"])
        old = (_f32(self._st0["velocity_request"]["x"]), _f32(self._st0["velocity_request"]["y"])) if had else None
        new = None if self.velocity_request is None else (_f32(self.velocity_request[0]), _f32(self.velocity_request[1]))
        if new != old:
            st["has_velocity_request"] = 0 if new is None else 1
            if new is not None:
                st["velocity_request"]["x"], st["velocity_request"]["y"] = new
            mask |= A.BODY_VELOCITY_REQUEST
        if self.body_type != int(self._st0["body_type"]):
            st["body_type"] = self.body_type
            mask |= A.BODY_TYPE
        if self.user_data != (int(self._st0["user_data_lo"]) | (int(self._st0["user_data_hi"]) << 64)):
            st["user_data_lo"], st["user_data_hi"] = self.user_data & 0xFFFFFFFFFFFFFFFF, self.user_data >> 64
            mask |= A.BODY_USER_DATA
        if mask:
            self._w.body_set(self._handle, st, mask)
            self._st0 = st
        return mask

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *_):
        if exc_type is None:
            self.commit()
        return False


def _body_desc(rbd):
    d = A.body_descs(1)
    d["position"]["x"], d["position"]["y"] = rbd.position
    d["position_old"]["x"], d["position_old"]["y"] = rbd.position_old
    d["gravity_mod"] = rbd.gravity_mod
    d["rotation"] = rbd.rotation
    d["scale"]["x"], d["scale"]["y"] = rbd.scale
    d["acceleration"]["x"], d["acceleration"]["y"] = rbd.acceleration
    if rbd.velocity_request is not None:
        d["has_velocity_request"] = 1
        d["velocity_request"]["x"], d["velocity_request"]["y"] = rbd.velocity_request
    d["calculated_velocity"]["x"], d["calculated_velocity"]["y"] = rbd.calculated_velocity
    d["user_data_lo"] = rbd.user_data & 0xFFFFFFFFFFFFFFFF
    d["user_data_hi"] = rbd.user_data >> 64
    d["body_type"] = rbd.body_type
    return d


def _collider_desc(c):
    d = A.collider_descs(1)
    for name, a in (("offset", c.offset), ("absolute_transform", c.absolute_transform)):
        d[name]["x_axis"]["x"], d[name]["x_axis"]["y"] = a.x_axis
        d[name]["y_axis"]["x"], d[name]["y_axis"]["y"] = a.y_axis
        d[name]["translation"]["x"], d[name]["translation"]["y"] = a.translation
    d["radius"] = c.radius
    d["shape_radius"] = c.radius
    if c.mass_override is not None:
        d["has_mass_override"] = 1
        d["mass_override"] = c.mass_override
    d["is_sensor"] = int(c.flags.is_sensor)
    d["memberships"] = c.collision_groups.memberships
    d["filter"] = c.collision_groups.filter
    d["user_data_lo"] = c.user_data & 0xFFFFFFFFFFFFFFFF
    d["user_data_hi"] = c.user_data >> 64
    return d


class Physics:
    """blobs::Physics (physics.rs:3-34). Handles are thunderdome Index bits (ints)."""

    def __init__(self, gravity=(0.0, 0.0), use_spatial_hash=False, device=-1, event_capacity=1 << 20):
        self._w = World(gravity=gravity, use_spatial_hash=use_spatial_hash, device=device)
        self._w.record_contacts(A.RECORD_EVENTS, event_capacity)  # the reference always feeds collision_send (physics.rs:304-311)
        self._events = []

    # ---- pub fields (physics.rs:6-33)
    def _prop(pid, cast=float):
        return property(lambda self: cast(self._w.get_param(pid)), lambda self, v: self._w.set_param(pid, v))

    substeps = _prop(A.PARAM_SUBSTEPS, int)
    joint_iterations = _prop(A.PARAM_JOINT_ITERATIONS, int)
    use_spatial_hash = _prop(A.PARAM_USE_SPATIAL_HASH, lambda v: bool(int(v)))
    collisions_enabled = _prop(A.PARAM_COLLISIONS_ENABLED, lambda v: bool(int(v)))
    accumulator = _prop(A.PARAM_ACCUMULATOR)
    time = _prop(A.PARAM_TIME)
    old_dt = _prop(A.PARAM_OLD_DT)

    @property
    def gravity(self):
        return (self._w.get_param(A.PARAM_GRAVITY_X), self._w.get_param(A.PARAM_GRAVITY_Y))

    @gravity.setter
    def gravity(self, g):
        self._w.set_param(A.PARAM_GRAVITY_X, g[0])
        self._w.set_param(A.PARAM_GRAVITY_Y, g[1])

    def constraints_push(self, c):
        self._w.constraint_push(c.position, c.radius)

    def constraints_clear(self):
        self._w.constraint_clear()

    # ---- methods
    def reset(self):  # physics.rs:71-76
        self._w.reset()

    def _pump(self):
        for e in self._w.events_drain():
            self._events.append(CollisionEvent(int(e["col_handle_a"]), int(e["col_handle_b"]),
                                               (float(e["impact_vel_a"]["x"]), float(e["impact_vel_a"]["y"])),
                                               (float(e["impact_vel_b"]["x"]), float(e["impact_vel_b"]["y"]))))

    def step(self, delta):  # physics.rs:78-82; panics become RuntimeError with the reference's message
        st = self._w.step(delta)
        self._pump()
        return st

    def fixed_step(self, frame_time):  # physics.rs:84-99
        st = self._w.fixed_step(frame_time)
        self._pump()
        return st

    def collision_recv(self):
        """Drains the CollisionEvent channel (demo/src/demos/balls.rs:104-106)."""
        ev, self._events = self._events, []
        return ev

    def insert_rbd(self, rbd):  # physics.rs:121-128
        return int(self._w.insert_bodies(_body_desc(rbd))[0])

    def insert_collider_with_parent(self, collider, rbd_handle):  # physics.rs:130-149
        return int(self._w.insert_colliders(_collider_desc(collider), np.array([rbd_handle], dtype=np.uint64))[0])

    def remove_rbd(self, handle):  # physics.rs:163-172 (a missing body only pushes an event in the reference)
        try:
            self._w.remove_body(handle)
        except BlobsError:
            pass

    def remove_col(self, handle):  # physics.rs:159-161
        try:
            self._w.remove_collider(handle)
        except BlobsError:
            pass

    def rbd_count(self):  # physics.rs:113-115
        return self._w.body_count()

    def get_rbd(self, handle):  # physics.rs:105-107 -> Option
        try:
            return self._w.body_get(handle)
        except BlobsError:
            return None

    def get_rbd_data(self, handle):  # physics.rs:101-103, RigidBodyData rigid_body.rs:18-26
        s = self.get_rbd(handle)
        if s is None:
            return None
        return {"position": (float(s["position"]["x"]), float(s["position"]["y"])),
                "velocity": (float(s["calculated_velocity"]["x"]), float(s["calculated_velocity"]["y"])),
                "angular_velocity": float(s["angular_velocity"]), "rotation": float(s["rotation"]),
                "center_of_mass": (float(s["center_of_mass"]["x"]), float(s["center_of_mass"]["y"])), "mass": float(s["calculated_mass"])}

    def get_mut_rbd(self, handle):  # physics.rs:109-111 -> Option<&mut RigidBody>: a host mirror with dirty tracking
        s = self.get_rbd(handle)
        return None if s is None else RigidBodyMut(self._w, handle, s)

    def set_rbd(self, handle, state, mask):  # get_mut_rbd physics.rs:109-111: write back the fields in `mask`
        self._w.body_set(handle, state, mask)

    def get_col(self, handle):  # physics.rs:117-119
        try:
            return self._w.collider_get(handle)
        except BlobsError:
            return None

    def rbd_position(self, handle):  # physics.rs:151-153
        s = self.get_rbd(handle)
        return None if s is None else (float(s["position"]["x"]), float(s["position"]["y"]))

    def col_position(self, handle):  # physics.rs:155-157
        s = self.get_col(handle)
        if s is None:
            return None
        t = s["desc"]["absolute_transform"]["translation"]
        return (float(t["x"]), float(t["y"]))

    def update_rigid_body_position(self, handle, offset):  # physics.rs:174-182
        self._w.body_translate(handle, offset)

    def create_fixed_joint(self, a, b, anchor_a=(0.0, 0.0), anchor_b=(0.0, 0.0)):  # physics.rs:184-207
        return self._w.joint_insert(a, b, anchor_a, anchor_b)

    def create_fixed_joint_with_distance(self, a, b, anchor_a, anchor_b, distance):  # physics.rs:209-239
        return self._w.joint_insert(a, b, anchor_a, anchor_b, distance)

    def springs_insert(self, spring):  # physics.springs.insert(Spring{..}) (demo/src/demos/joints.rs:59-66)
        return self._w.spring_insert(spring.rigid_body_a, spring.rigid_body_b, spring.rest_length, spring.stiffness, spring.damping)

    def debug_data(self):  # physics.rs:479-481 / debug.rs:34-91
        """DebugData { bodies, joints, colliders, springs } from one library call (blobs_debug_data)."""
        d = self._w.debug_data()

        def affine(row):
            return Affine2((float(row[0]), float(row[1])), (float(row[2]), float(row[3])), (float(row[4]), float(row[5])))

        return {
            "bodies": [affine(r) for r in d["bodies"]],                                                        # DebugRigidBody.transform
            "joints": [((float(r[0]), float(r[1])), (float(r[2]), float(r[3]))) for r in d["joints"]],         # DebugJoint {body_a, body_b}
            "colliders": [(affine(r), float(rad)) for r, rad in zip(d["colliders"], d["collider_radius"])],    # DebugCollider {transform, radius}
            "springs": [((float(r[0]), float(r[1])), (float(r[2]), float(r[3]))) for r in d["springs"]],       # DebugSpring {body_a, body_b}
        }
