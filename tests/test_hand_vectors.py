"""Hand-derived known-answer vectors for Physics::step (SURVEY.md §8c 'vectors the new repo must author').

Every expected value is computed here in numpy float32 following the reference formula it cites, independently
of both implementations. Each test runs against the CPU oracle (pins the oracle; part of the CPU suite) and,
under -m gpu, against the CUDA library through its C ABI (same vectors, bit-exact)."""
import numpy as np
import pytest

from .helpers import backend_params, make_backend, sphere

f32 = np.float32


@pytest.fixture(params=backend_params())
def be(request):
    return make_backend(request.param)


def test_handles_are_thunderdome_bits(be):
    w = be.make()
    b0, c0 = sphere(w, (0, 0))
    b1, c1 = sphere(w, (5, 0))
    assert b0 == 1 << 32 and b1 == (1 << 32) | 1  # first handle = 1<<32 (demo/src/demos/balls.rs:167)
    assert c0 == 1 << 32
    w.remove_body(b0)
    assert w.body_count() == 1 and w.collider_count() == 1
    b2, c2 = sphere(w, (9, 0))
    assert b2 == (2 << 32) | 0 and c2 == (2 << 32) | 0  # LIFO reuse, generation bumped


def test_mass_is_doubled(be):  # Q1: handle pushed twice => calculated_mass = 2 * (2r)
    w = be.make()
    b0, _ = sphere(w, (0, 0), r=0.5)
    st = w.body_get(b0)
    assert st["calculated_mass"] == f32(2.0)
    assert st["inertia"] == f32(2 * 0.5 * 1.0 * 0.25)
    assert len(w.body_colliders(b0)) == 2


def test_two_equal_spheres_overlap(be):
    """(1) r=0.5, centres 0.8 apart on x => overlap 0.2, equal mass => each moves 0.1 (ratio 0.5). One substep."""
    A = be.A
    w = be.make()
    w.set_param(A.PARAM_SUBSTEPS, 1)
    w.record_contacts(A.RECORD_PAIRS, 1024)
    sphere(w, (0.0, 0.0))
    sphere(w, (0.8, 0.0))
    r = w.step(1.0 / 60.0)
    assert r["collisions"] == 1
    dt = f32(f32(1.0 / 60.0) / f32(1))
    # a = later slot (body 1): axis = (0.8, 0); dist = 0.8; delta = 1 - 0.8; ratio = 1 - 2/(2+2) = 0.5
    dist = f32(0.8)
    delta = f32(f32(1.0) - dist)
    n = f32(dist / dist)
    push_a = f32(f32(f32(0.5) * delta) * n)
    push_b = f32(f32(f32(f32(1.0) - f32(0.5)) * delta) * n)
    xa = f32(f32(0.8) + push_a)
    xb = f32(f32(0.0) - push_b)
    # verlet with zero velocity/acc: the first dynamic body (slot 0) sees ratio dt/old_dt = dt/1.0 (Q2), slot 1 sees dt/dt
    da = f32(f32(xa - f32(0.8)) * f32(dt / dt))
    db = f32(f32(xb - f32(0.0)) * f32(dt / f32(1.0)))
    st, _ = w.download_bodies()
    assert st["position"]["x"][1] == f32(xa + f32(da + f32(0)))
    assert st["position"]["x"][0] == f32(xb + f32(db + f32(0)))
    assert st["position_old"]["x"][1] == xa and st["position_old"]["x"][0] == xb
    pairs = w.pairs_drain()
    assert len(pairs) == 1 and pairs[0].tolist() == [[1, 0]]


def test_unequal_radii_ratio(be):
    """(2) ratio from the doubled 2r masses: a (later slot) moves by ratio = 1 - m_a/(m_a+m_b)."""
    A = be.A
    w = be.make()
    w.set_param(A.PARAM_SUBSTEPS, 1)
    w.set_param(A.PARAM_OLD_DT, 1.0 / 60.0)  # neutralise Q2 for this vector
    sphere(w, (0.0, 0.0), r=0.25)
    sphere(w, (0.5, 0.0), r=0.5)
    w.step(1.0 / 60.0)
    ma, mb = f32(2 * f32(0.5) * 2), f32(2 * f32(0.25) * 2)
    ratio = f32(f32(1.0) - f32(ma / f32(ma + mb)))
    delta = f32(f32(f32(0.5) + f32(0.25)) - f32(0.5))
    pa = f32(f32(0.5) + f32(f32(ratio * delta) * f32(1.0)))
    pb = f32(f32(0.0) - f32(f32(f32(f32(1.0) - ratio) * delta) * f32(1.0)))
    st, _ = w.download_bodies()
    # velocity carries: pos += (pos - pos_old) * 1
    assert st["position"]["x"][1] == f32(pa + f32(pa - f32(0.5)))
    assert st["position"]["x"][0] == f32(pb + f32(pb - f32(0.0)))


def test_free_fall_verlet_with_old_dt_quirk(be):
    """(3)+(4) single free body, g=(0,-30), 8 substeps of dt=1/480, velocity_request (Q9). Q2: the first substep
    scales the displacement by dt/1.0."""
    A = be.A
    w = be.make(gravity=(0.0, -30.0))
    b0, _ = sphere(w, (1.0, 2.0), velocity_request=(3.0, 0.0))
    w.step(1.0 / 60.0)
    dt = f32(f32(1.0 / 60.0) / f32(8))
    px, py = f32(1.0), f32(2.0)
    pox, poy = f32(px - f32(f32(3.0) * dt)), f32(py - f32(f32(0.0) * dt))  # Q9
    old_dt = f32(1.0)
    for _ in range(8):
        ay = f32(f32(0.0) + f32(f32(-30.0) * f32(1.0)))
        ratio = f32(dt / old_dt)
        old_dt = dt
        dx, dy = f32(f32(px - pox) * ratio), f32(f32(py - poy) * ratio)
        pox, poy = px, py
        px = f32(px + f32(dx + f32(f32(f32(0.0) * dt) * dt)))
        py = f32(py + f32(dy + f32(f32(ay * dt) * dt)))
        vx, vy = f32(dx / dt), f32(dy / dt)
    st = w.body_get(b0)
    assert st["position"]["x"] == px and st["position"]["y"] == py
    assert st["position_old"]["x"] == pox and st["position_old"]["y"] == poy
    assert st["calculated_velocity"]["x"] == vx and st["calculated_velocity"]["y"] == vy
    assert st["has_velocity_request"] == 0
    assert f32(w.get_param(A.PARAM_OLD_DT)) == dt
    assert w.get_param(A.PARAM_TIME) == 1.0 / 60.0


def test_spring_componentwise_force(be):
    """(5) one spring; Q7: force = dir * (k*(L-rest) - damp_vec) component-wise."""
    A = be.A
    w = be.make()
    w.set_param(A.PARAM_SUBSTEPS, 1)
    w.set_param(A.PARAM_OLD_DT, 1.0 / 60.0)
    w.set_param(A.PARAM_COLLISIONS_ENABLED, 0)
    a, _ = sphere(w, (0.0, 0.0))
    b, _ = sphere(w, (3.0, 4.0), calculated_velocity=(1.0, -2.0))
    w.spring_insert(a, b, 4.0, 10.0, 0.5)
    w.step(1.0 / 60.0)
    dt = f32(1.0 / 60.0)
    d = (f32(3.0), f32(4.0))
    L = f32(np.sqrt(f32(f32(d[0] * d[0]) + f32(d[1] * d[1]))))
    u = (f32(d[0] / L), f32(d[1] / L))
    rv = (f32(f32(0) - f32(1.0)), f32(f32(0) - f32(-2.0)))
    dd = f32(f32(0.5) * f32(f32(rv[0] * u[0]) + f32(rv[1] * u[1])))
    damp = (f32(dd * u[0]), f32(dd * u[1]))
    s = f32(f32(10.0) * f32(L - f32(4.0)))
    fm = (f32(s - damp[0]), f32(s - damp[1]))
    F = (f32(u[0] * fm[0]), f32(u[1] * fm[1]))
    m = f32(2.0)
    acc_a = (f32(f32(0) + f32(F[0] / m)), f32(f32(0) + f32(F[1] / m)))
    acc_b = (f32(f32(0) + f32(f32(-F[0]) / m)), f32(f32(0) + f32(f32(-F[1]) / m)))
    st, _ = w.download_bodies()
    assert st["position"]["x"][0] == f32(f32(0.0) + f32(f32(0.0) + f32(f32(acc_a[0] * dt) * dt)))
    assert st["position"]["y"][0] == f32(f32(0.0) + f32(f32(0.0) + f32(f32(acc_a[1] * dt) * dt)))
    assert st["position"]["x"][1] == f32(f32(3.0) + f32(f32(0.0) + f32(f32(acc_b[0] * dt) * dt)))
    assert st["position"]["y"][1] == f32(f32(4.0) + f32(f32(0.0) + f32(f32(acc_b[1] * dt) * dt)))


@pytest.mark.parametrize("static_a", [False, True])
def test_joint_dynamic_dynamic_and_static(be, static_a):
    """(6) one joint; Q8: static branch scaled by inv_mass_sum, anchors unrotated, angle terms."""
    A = be.A
    w = be.make()
    w.set_param(A.PARAM_SUBSTEPS, 1)
    w.set_param(A.PARAM_JOINT_ITERATIONS, 1)
    w.set_param(A.PARAM_OLD_DT, 1.0 / 60.0)
    w.set_param(A.PARAM_COLLISIONS_ENABLED, 0)
    a, _ = sphere(w, (0.0, 0.0), r=0.5, body_type=A.BODY_STATIC if static_a else A.BODY_DYNAMIC)
    b, _ = sphere(w, (2.0, 0.0), r=0.25)
    w.joint_insert(a, b, (0.0, 0.0), (0.0, 0.0), 1.0)
    w.step(1.0 / 60.0)
    ma, mb = f32(2.0), f32(1.0)
    dist = f32(2.0)
    corr = f32(f32(f32(dist - f32(1.0)) * f32(2.0)) / dist)
    ims = f32(f32(f32(1.0) / ma) + f32(f32(1.0) / mb))
    st, _ = w.download_bodies()
    if static_a:
        pb = f32(f32(2.0) - f32(ims * corr))
        assert st["position"]["x"][0] == f32(0.0)
        assert st["position"]["x"][1] == f32(pb + f32(pb - f32(2.0)))
    else:
        ratio = f32(f32(f32(1.0) / ma) / ims)
        pa = f32(f32(0.0) + f32(ratio * corr))
        pb = f32(f32(2.0) - f32(f32(f32(1.0) - ratio) * corr))
        assert st["position"]["x"][0] == f32(pa + f32(pa - f32(0.0)))
        assert st["position"]["x"][1] == f32(pb + f32(pb - f32(2.0)))
    # d = (2,0): angle_a = atan2(0,2) = 0; angle_b = -atan2(0,-2) = -pi; rc = (-pi - 0 - 0) * 0.5
    rc = f32(f32(f32(-np.arctan2(f32(0.0), f32(-2.0))) - f32(0.0)) * f32(0.5))
    dt = f32(1.0 / 60.0)
    tol = 0 if be.name == "oracle" else 2e-7  # GPU atan2f may differ from libm by ulps
    assert abs(st["rotation"][0] - f32(f32(0.0) + f32(rc * dt))) <= tol
    assert abs(st["rotation"][1] - f32(f32(0.0) - f32(rc * dt))) <= tol


def test_joint_distance_from_positions(be):
    """create_fixed_joint (physics.rs:184-207): distance = |pos_a + anchor_a - pos_b - anchor_b|, so a joint created at
    rest produces no positional correction."""
    A = be.A
    w = be.make()
    w.set_param(A.PARAM_SUBSTEPS, 1)
    w.set_param(A.PARAM_OLD_DT, 1.0 / 60.0)
    w.set_param(A.PARAM_COLLISIONS_ENABLED, 0)
    a, _ = sphere(w, (0.25, 1.0))
    b, _ = sphere(w, (3.0, -2.0))
    w.joint_insert(a, b, (0.5, 0.0), (0.0, 0.25))
    w.step(1.0 / 60.0)
    st, _ = w.download_bodies()
    assert st["position"]["x"].tolist() == [0.25, 3.0] and st["position"]["y"].tolist() == [1.0, -2.0]


def test_circle_clamp_ignores_radius_and_includes_static(be):
    """(7) Q11; snapshot is taken before the clamp (Q3)."""
    A = be.A
    w = be.make()
    w.set_param(A.PARAM_SUBSTEPS, 1)
    w.constraint_push((0.0, 0.0), 4.0)
    sphere(w, (3.0, 4.0))
    sphere(w, (-6.0, 0.0), body_type=A.BODY_STATIC)
    w.step(1.0 / 60.0)
    st, _ = w.download_bodies()
    L = f32(5.0)
    assert st["position"]["x"][0] == f32(f32(0.0) + f32(f32(f32(3.0) / L) * f32(4.0)))
    assert st["position"]["y"][0] == f32(f32(0.0) + f32(f32(f32(4.0) / L) * f32(4.0)))
    assert st["position"]["x"][1] == f32(-4.0)
    cs, _ = w.download_colliders()
    assert cs["desc"]["absolute_transform"]["translation"]["x"][0] == f32(3.0)
    assert cs["desc"]["absolute_transform"]["translation"]["x"][1] == f32(-6.0)


def test_sensor_pair_counts_but_does_not_push(be):
    """(8) Q6: sensors are counted and reported but never pushed; they do not contribute mass."""
    A = be.A
    w = be.make()
    w.set_param(A.PARAM_SUBSTEPS, 1)
    w.record_contacts(A.RECORD_EVENTS, 1024)
    sphere(w, (0.0, 0.0))
    sphere(w, (0.5, 0.0), col={"is_sensor": 1})
    r = w.step(1.0 / 60.0)
    assert r["collisions"] == 1
    st, _ = w.download_bodies()
    assert st["position"]["x"].tolist() == [0.0, 0.5]
    assert st["calculated_mass"][1] == f32(1.0)  # sensor excluded from mass => 0 => 1.0
    ev = w.events_drain()
    assert len(ev) == 1 and ev[0]["col_handle_a"] == (1 << 32) | 1 and ev[0]["col_handle_b"] == 1 << 32


def test_groups_none_never_collides(be):
    """(9) groups(0,0) (demo/src/demos/balls.rs:22)."""
    w = be.make()
    sphere(w, (0.0, 0.0))
    sphere(w, (0.5, 0.0), col={"memberships": 0, "filter": 0})
    assert w.step(1.0 / 60.0)["collisions"] == 0


def test_static_body_is_pushed_by_contacts(be):
    """Q5: no static check in the contact push (physics.rs:298-299)."""
    A = be.A
    w = be.make()
    w.set_param(A.PARAM_SUBSTEPS, 1)
    sphere(w, (0.0, 0.0), body_type=A.BODY_STATIC)
    sphere(w, (0.8, 0.0))
    w.step(1.0 / 60.0)
    st, _ = w.download_bodies()
    assert st["position"]["x"][0] < 0.0
    assert st["position_old"]["x"][0] == st["position"]["x"][0]
    assert st["calculated_velocity"]["x"][0] == 0.0


def test_first_contact_pass_uses_caller_snapshot(be):
    """Q3: with the builder-default absolute_transform (IDENTITY) new colliders sit at the origin in their first pass,
    which is the coincident-centre branch (physics.rs:272-286): +-0.01 push-out, then a normal contact at distance 0.02."""
    A = be.A
    w = be.make()
    w.set_param(A.PARAM_SUBSTEPS, 1)
    w.set_param(A.PARAM_OLD_DT, 1.0 / 60.0)
    b = A.body_descs(2)
    b["position"]["x"] = [0.0, 10.0]
    b["position_old"] = b["position"]
    bh = w.insert_bodies(b)
    w.insert_colliders(A.collider_descs(2), bh)  # both snapshots at (0,0)
    r = w.step(1.0 / 60.0)
    assert r["collisions"] == 1 and r["coincident_pairs"] == 1
    # a = slot 1: pos 10 + 0.01, snapshot a = 10.01 (+offset 0); b: -0.01; axis = 10.02 > min_dist => delta negative
    pa = f32(f32(10.0) + f32(0.01))
    pb = f32(f32(0.0) - f32(0.01))
    axis = f32(pa - pb)
    delta = f32(f32(1.0) - axis)
    n = f32(axis / axis)
    pa2 = f32(pa + f32(f32(f32(0.5) * delta) * n))
    pb2 = f32(pb - f32(f32(f32(0.5) * delta) * n))
    st, _ = w.download_bodies()
    if be.name == "oracle":
        assert st["position_old"]["x"][1] == pa2 and st["position_old"]["x"][0] == pb2
    else:
        # documented divergence: the GPU pass is Jacobi on the snapshot, so the re-measured axis uses the
        # snapshots (0.02 apart), not the live body positions (10.02 apart)
        axis_g = f32(f32(f32(0.0) + f32(0.01)) - f32(f32(0.0) - f32(0.01)))
        delta_g = f32(f32(1.0) - axis_g)
        assert st["position_old"]["x"][1] == f32(pa + f32(f32(f32(0.5) * delta_g) * f32(axis_g / axis_g)))


def test_fixed_step_accumulator(be):
    """Q12: at most 3 integrate calls per fixed_step, leftover retained."""
    A = be.A
    w = be.make()
    sphere(w, (0, 0))
    assert w.fixed_step(0.01)["steps_run"] == 0
    assert w.fixed_step(0.01)["steps_run"] == 1
    assert w.fixed_step(1.0)["steps_run"] == 3
    assert w.get_param(A.PARAM_ACCUMULATOR) == pytest.approx(0.02 - 1 / 60 + 1.0 - 3 / 60)
    assert w.get_param(A.PARAM_TIME) == pytest.approx(4 / 60)


def test_use_spatial_hash_panics(be):
    """physics.rs:410-415: stepping with use_spatial_hash panics / returns BLOBS_ERR_SPATIAL_HASH."""
    w = be.make(use_spatial_hash=True)
    sphere(w, (0, 0))
    with pytest.raises(RuntimeError, match="spatial collisions not supported"):
        w.step(1 / 60)


def test_apply_force_and_gravity_mod(be):
    """RigidBody::apply_force (rigid_body.rs:155-160) is consumed by the first substep only; gravity_mod scales gravity."""
    A = be.A
    w = be.make(gravity=(0.0, -10.0))
    w.set_param(A.PARAM_SUBSTEPS, 2)
    w.set_param(A.PARAM_OLD_DT, 1.0 / 120.0)
    b0, _ = sphere(w, (0.0, 0.0), gravity_mod=0.5)
    w.body_apply_force(b0, (4.0, 0.0))
    w.step(1.0 / 60.0)
    dt = f32(f32(1.0 / 60.0) / f32(2))
    m = f32(2.0)
    ax0 = f32(f32(0.0) + f32(f32(4.0) / m))
    gy = f32(f32(-10.0) * f32(0.5))
    px = py = pox = poy = f32(0.0)
    for sub in range(2):
        ax = f32((ax0 if sub == 0 else f32(0.0)) + f32(f32(0.0) * f32(0.5)))
        ay = f32(f32(0.0) + gy)
        dx, dy = f32(f32(px - pox) * f32(dt / dt)), f32(f32(py - poy) * f32(dt / dt))
        pox, poy = px, py
        px = f32(px + f32(dx + f32(f32(ax * dt) * dt)))
        py = f32(py + f32(dy + f32(f32(ay * dt) * dt)))
    st = w.body_get(b0)
    assert st["position"]["x"] == px and st["position"]["y"] == py
    assert st["acceleration"]["x"] == 0.0 and st["acceleration"]["y"] == 0.0


def test_remove_collider_removes_orphan_body(be):
    """collider.rs:143-158: removing the last collider removes the parent body too."""
    w = be.make()
    b0, c0 = sphere(w, (0, 0))
    b1, c1 = sphere(w, (5, 0))
    w.remove_collider(c0)
    assert w.body_count() == 1 and w.collider_count() == 1
    w.step(1 / 60)
    st, hd = w.download_bodies()
    assert hd[0] == 0 and hd[1] == b1
