"""RigidBody's force / impulse / velocity API (rigid_body.rs:130-214) through `get_mut_rbd` (physics.rs:109-111): the host
mirror with dirty tracking (blobs_b200.physics.RigidBodyMut), plus the remove_* semantics around it - auto-removal of a body
whose last collider goes (collider.rs:134-164) and the two soft-error messages of the event ring (events.rs, rigid_body.rs:266-275).
Known answers are computed here in numpy float32 from the cited formulas; the same script then runs on the CPU oracle and on
the CUDA library (-m gpu; -m emu = host-compiled kernels) and the worlds must agree bit for bit after stepping."""
import numpy as np
import pytest

from blobs_b200 import _abi as A
from blobs_b200.physics import RigidBodyMut, RigidBodyType

from .helpers import assert_bodies_bit_equal, backend_params, make_backend, sphere

f32 = np.float32


def _mut(w, h):
    return RigidBodyMut(w, h, w.body_get(h))


def _script(w):
    """three balls: a dynamic one that gets impulses / forces, one that gets a velocity, a static one that ignores everything"""
    a, _ = sphere(w, (0.0, 0.0), r=0.5)                      # calculated_mass = 2 * (2 r) = 2
    b, _ = sphere(w, (3.0, 0.0), r=0.25)                     # mass 1
    s, _ = sphere(w, (6.0, 0.0), r=0.5, body_type=A.BODY_STATIC)
    w.step(1.0 / 60.0)                                        # gives every body a calculated_velocity (zero here) and snapshots
    masks = {}
    with _mut(w, a) as m:
        v0 = m.get_velocity()                                 # calculated_velocity: zero (no gravity, nothing moved)
        m.apply_impulse((2.0, -1.0))                          # add_velocity(J / m): velocity_request = v0 + (1, -0.5)
        m.apply_force((0.5, 4.0))                             # acceleration += F / m = (0.25, 2)
        m.apply_force_at_point((0.0, 2.0), (1.0, 0.0))        # + (0, 1); torque += lever.perp_dot(F) = 1*2 - 0*0 = 2
        m.apply_torque_at_point((3.0, 0.0), (0.0, -2.0))      # torque += 0*0 - (-2)*3 = 6
        # add_velocity reads calculated_velocity, not the pending request (rigid_body.rs:151-153,186-188): the second impulse
        # REPLACES the first one's request with v0 + (0, 0.5); angular_velocity += (0.5*1 - 0) / inertia
        m.apply_impulse_at_point((0.0, 1.0), (0.5, 0.0))
        m.accelerate((1.0, 0.0))                              # acceleration += (1, 0)
        want_a = dict(velocity_request=(v0[0] + f32(0.0), v0[1] + f32(0.5)), acceleration=(f32(0.25) + f32(1.0), f32(2.0) + f32(1.0)),
                      torque=f32(8.0), angular_velocity=f32(0.5) / m.inertia)
    masks["a"] = m.commit()                                   # second commit: nothing left to write
    with _mut(w, b) as m:
        m.set_velocity((0.0, 3.0))
        m.gravity_mod = f32(0.5)
        m.position = (m.position[0], m.position[1] + f32(0.125))
    with _mut(w, s) as m:
        assert m.is_static() and not m.is_dynamic() and not m.is_kinematic()
        m.apply_impulse((5.0, 5.0)); m.apply_force((5.0, 5.0)); m.apply_force_at_point((1.0, 1.0), (9.0, 9.0))
        m.apply_torque_at_point((1.0, 1.0), (9.0, 9.0)); m.apply_impulse_at_point((1.0, 1.0), (9.0, 9.0))
    masks["s"] = m.commit()
    return (a, b, s), want_a, masks


@pytest.mark.parametrize("backend", backend_params())
def test_force_impulse_velocity_api_through_get_mut_rbd(backend):
    be = make_backend(backend)
    w = be.make(gravity=(0.0, 0.0))
    (a, b, s), want_a, masks = _script(w)
    assert masks == {"a": 0, "s": 0}                          # clean mirror / static body: no write-back at all
    sa, sb = w.body_get(a), w.body_get(b)
    assert sa["has_velocity_request"] == 1
    assert (f32(sa["velocity_request"]["x"]), f32(sa["velocity_request"]["y"])) == want_a["velocity_request"]
    assert (f32(sa["acceleration"]["x"]), f32(sa["acceleration"]["y"])) == want_a["acceleration"]
    assert f32(sa["torque"]) == want_a["torque"] and f32(sa["angular_velocity"]) == want_a["angular_velocity"]
    assert sb["has_velocity_request"] == 1 and f32(sb["velocity_request"]["y"]) == f32(3.0)
    assert f32(sb["gravity_mod"]) == f32(0.5) and f32(sb["position"]["y"]) == f32(0.125)
    if backend == "oracle":
        return
    # the same script on the oracle, then both worlds step: the staged writes must have landed exactly where the
    # reference's &mut RigidBody would have put them
    ow = make_backend("oracle").make(gravity=(0.0, 0.0))
    _script(ow)
    for _ in range(3):
        w.step(1.0 / 60.0)
        ow.step(1.0 / 60.0)
    got, _ = w.download_bodies()
    want, _ = ow.download_bodies()
    assert_bodies_bit_equal(got, want)
    assert np.allclose(got["rotation"], want["rotation"], rtol=1e-6, atol=1e-7)
    assert np.array_equal(got["angular_velocity"], want["angular_velocity"]) and np.array_equal(got["torque"], want["torque"])


def check_removal_semantics_and_event_ring():
    from blobs_b200 import events
    from blobs_b200.physics import Affine2, ColliderBuilder, Physics, RigidBodyBuilder

    events.clear_event_history()
    physics = Physics(gravity=(0.0, 0.0))
    rbd = physics.insert_rbd(RigidBodyBuilder().position((1.5, -2.0)).build())
    c1 = physics.insert_collider_with_parent(ColliderBuilder().radius(0.5).absolute_transform(Affine2.from_translation((1.5, -2.0))).build(), rbd)
    c2 = physics.insert_collider_with_parent(ColliderBuilder().radius(0.25).build(), rbd)
    assert physics.get_rbd_data(rbd)["mass"] == pytest.approx(2 * (1.0 + 0.5))     # doubled sum of 2r (Q1)
    physics.remove_col(c1)                                                         # collider.rs:134-164: retain + update mass
    assert physics.rbd_count() == 1 and physics.get_col(c1) is None
    assert physics.get_rbd_data(rbd)["mass"] == pytest.approx(2 * 0.5)
    assert events.event_history() == []
    physics.remove_col(c2)                                                         # last collider: the body goes too
    assert physics.rbd_count() == 0 and physics.get_rbd(rbd) is None and physics.get_mut_rbd(rbd) is None
    physics.remove_rbd(rbd)                                                        # rigid_body.rs:266-275: logged, no panic
    hist = events.event_history()
    assert [(e.message, e.severity) for e in hist] == [("rbd removed because colliders.len() == 0", events.Severity.Info),
                                                       ("removing a non-existent rigid body", events.Severity.Error)]
    assert hist[0].position == (1.5, -2.0) and hist[0].col_handle == c2 and hist[0].rbd_handle == rbd
    assert hist[1].position is None and hist[1].col_handle is None and hist[1].rbd_handle == rbd
    assert hist[0].time_data == (0.0, 0.0)                                         # TimeData is never advanced (events.rs:11-18)
    for _ in range(1005):                                                          # ring: at most 1000 entries (events.rs:35-38)
        physics.remove_rbd(rbd)
    hist = events.event_history()
    assert len(hist) == 1000 and all(e.message == "removing a non-existent rigid body" for e in hist)
    events.clear_event_history()
    # the slot is reused LIFO with a bumped generation (thunderdome)
    again = physics.insert_rbd(RigidBodyBuilder().build())
    assert again & 0xFFFFFFFF == rbd & 0xFFFFFFFF and again >> 32 == (rbd >> 32) + 1


@pytest.mark.gpu
def test_removal_semantics_and_event_ring():
    check_removal_semantics_and_event_ring()
