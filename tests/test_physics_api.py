"""The reference-shaped host API (blobs_b200.physics mirrors blobs::Physics, the builders and the handle types) driven the way
the reference's demo drives it (demo/src/main.rs:51-54, demo/src/simulation.rs:122-146, demo/src/demos/balls.rs:99-106)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _spawn(physics, pos, radius, vel=None):
    from blobs_b200.physics import Affine2, ColliderBuilder, RigidBodyBuilder

    b = RigidBodyBuilder().position(pos)
    if vel is not None:
        b = b.velocity_request(vel)
    rbd = physics.insert_rbd(b.build())
    col = physics.insert_collider_with_parent(ColliderBuilder().radius(radius).absolute_transform(Affine2.from_translation(pos)).build(), rbd)
    return rbd, col


def test_balls_demo_flow():
    from blobs_b200.physics import Constraint, Physics

    physics = Physics(gravity=(0.0, -30.0), use_spatial_hash=False)
    assert physics.substeps == 8 and physics.joint_iterations == 4 and physics.collisions_enabled  # physics.rs:46-47,62
    physics.constraints_push(Constraint(position=(0.0, 0.0), radius=4.0))
    handles = [_spawn(physics, (0.3 * (i % 8) - 1.0, 0.3 * (i // 8)), 0.1 + 0.01 * (i % 5), vel=(0.5, -1.0)) for i in range(64)]
    assert handles[0][0] == 1 << 32 and physics.rbd_count() == 64
    n_events = 0
    for _ in range(90):
        physics.fixed_step(1 / 60)
        for ev in physics.collision_recv():
            assert ev.col_handle_a & 0xFFFFFFFF > ev.col_handle_b & 0xFFFFFFFF  # a = later slot (physics.rs:248-249)
            n_events += 1
    assert n_events > 0
    assert physics.time == pytest.approx(90 / 60, rel=1e-9)
    for rbd, col in handles:
        x, y = physics.rbd_position(rbd)
        assert math.hypot(x, y) <= 4.0 + 1e-4          # circle constraint (physics.rs:377-395)
        assert physics.col_position(col) is not None
    d = physics.get_rbd_data(handles[3][0])
    assert d["mass"] == pytest.approx(2 * 2 * (0.1 + 0.03))   # doubled 2r
    dbg = physics.debug_data()
    assert len(dbg["bodies"]) == 64 and len(dbg["colliders"]) == 64
    physics.remove_rbd(handles[0][0])
    assert physics.rbd_count() == 63 and physics.get_rbd(handles[0][0]) is None and physics.get_col(handles[0][1]) is None
    physics.remove_rbd(handles[0][0])   # removing twice only logs an event in the reference
    physics.reset()
    assert physics.rbd_count() == 0


def test_joints_springs_and_panics():
    from blobs_b200.physics import Physics, Spring

    physics = Physics(gravity=(0.0, 0.0))
    a, _ = _spawn(physics, (0.0, 0.0), 0.2)
    b, _ = _spawn(physics, (1.0, 0.0), 0.2)
    c, _ = _spawn(physics, (3.0, 0.0), 0.2)
    physics.create_fixed_joint(a, b)
    physics.springs_insert(Spring(b, c, rest_length=1.0, stiffness=50.0, damping=1.0))
    for _ in range(30):
        physics.step(1 / 60)
    pa, pb, pc = (np.array(physics.rbd_position(h)) for h in (a, b, c))
    assert abs(np.linalg.norm(pb - pa) - 1.0) < 0.05          # joint keeps its creation distance
    assert np.linalg.norm(pc - pb) < 2.0                       # spring pulls c towards b
    with pytest.raises(RuntimeError, match="identical indices"):
        physics.create_fixed_joint(a, a)                       # thunderdome get2_mut panic (physics.rs:191-196)
    physics.remove_rbd(c)
    with pytest.raises(RuntimeError, match="removed rigid body"):
        physics.step(1 / 60)                                    # dangling spring: unwrap() panic (springs.rs:26-29)
    p2 = Physics(gravity=(0.0, 0.0), use_spatial_hash=True)
    _spawn(p2, (0.0, 0.0), 0.2)
    with pytest.raises(RuntimeError, match="spatial collisions not supported right now"):
        p2.step(1 / 60)                                         # physics.rs:410-415
    with pytest.raises(RuntimeError, match="parent rigid body must exist"):
        from blobs_b200.physics import ColliderBuilder
        p2.insert_collider_with_parent(ColliderBuilder().build(), 12345 << 32)   # physics.rs:142
