"""Scene-level parity: the CUDA library (through its C ABI) against the CPU oracle on identical seeded scenes.

Bars (BASELINE.json north_star / SURVEY §8d): contact pair sets per substep — exact set equality; cell coords — exact;
positions / velocities — bit-exact in ordered mode (the GPU applies each body's contributions in the reference's
pair-loop order); rotation of jointed bodies — 1e-5 relative (GPU atan2f/sincosf differ from libm by ulps)."""
import os

import numpy as np
import pytest

from blobs_b200 import _abi as A
from blobs_b200 import scenes as S

from .helpers import assert_bodies_bit_equal, bits, sphere

pytestmark = pytest.mark.gpu

_PIPELINES = {
    "grid-per-lane": {A.PARAM_LIST: 0, A.PARAM_POOL: 0},
    "grid-cooperative-crowded": {A.PARAM_LIST: 0, A.PARAM_POOL: 1, A.PARAM_CROWDED: 1},
    "lists": {A.PARAM_LIST: 1},
    "lists-tiny-skin": {A.PARAM_LIST: 1, A.PARAM_SKIN: 0.02},   # rebuilt almost every substep
    "lists-crowded": {A.PARAM_LIST: 1, A.PARAM_CROWDED: 1},
}




def _pair(gravity, scene, grid_oracle=False, **gpu_kw):
    import blobs_b200
    from oracle import oracle_py

    g = blobs_b200.World(gravity=gravity, **gpu_kw)
    o = oracle_py.OracleWorld(gravity=gravity, grid_pairs=grid_oracle, maintain_spatial_hash=False, record_events=False)
    hg = S.build(g, scene)
    ho = S.build(o, scene)
    for k in ("bodies", "colliders"):
        assert np.array_equal(hg[k], ho[k]), "handle assignment must match thunderdome's"
    return g, o


def _compare_step(g, o, fields=("position", "position_old", "calculated_velocity", "acceleration")):
    sg, hg = g.download_bodies()
    so, ho = o.download_bodies()
    assert np.array_equal(hg, ho)
    assert_bodies_bit_equal(sg, so, fields)
    cg, _ = g.download_colliders()
    co, _ = o.download_colliders()
    for c in ("x", "y"):
        assert np.array_equal(bits(cg["desc"]["absolute_transform"]["translation"][c]), bits(co["desc"]["absolute_transform"]["translation"][c])), "snapshot"


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_cfg1_pairs_and_positions_every_substep(seed):
    """cfg1 (1024 spheres, r~U[0.05,0.2), circle container) vs the brute-force oracle: pair set each substep, state each step."""
    sc = S.cfg1(seed)
    g, o = _pair(sc.gravity, sc)
    g.record_contacts(A.RECORD_PAIRS, 1 << 20)
    n_pairs = 0
    for step in range(40):
        rg = g.step(1 / 60)
        o.step(1 / 60)
        pg, po = g.pairs_drain(), o.pairs_drain()
        assert len(pg) == len(po) == 8
        for s, (x, y) in enumerate(zip(pg, po)):
            assert np.array_equal(x, y), f"pair set differs at step {step} substep {s}: gpu {len(x)} oracle {len(y)}"
            n_pairs += len(x)
        assert rg["events_dropped"] == 0 and rg["nan_detected"] == 0
        _compare_step(g, o)
    assert n_pairs > 1000, "scene must actually collide"
    assert o.coincident_total() == 0
    cxg, cyg = g.cell_coords()
    cxo, cyo = o.cell_coords()
    assert np.array_equal(cxg, cxo) and np.array_equal(cyg, cyo)
    assert g.kernel_info()["fused_path"] == 1


def test_cfg1_long_run_600_steps():
    """600 steps of cfg1: still bit-identical to the reference path at the end; collision counter agrees."""
    sc = S.cfg1(1)
    g, o = _pair(sc.gravity, sc)
    tot = 0
    for _ in range(6):
        tot += g.step(1 / 60, n=100)["collisions"]
        o.step(1 / 60, n=100)
    _compare_step(g, o)
    assert tot == o.step(1 / 60, n=0)["collisions"]


@pytest.mark.parametrize("fused", [1, 0])
@pytest.mark.parametrize("varied", [False, True])
def test_dense_pile_vs_grid_oracle(fused, varied):
    """16k overlapping spheres squeezed by a tight circle (busy contacts + clamps) vs the cell-list oracle, fused and
    split pipelines."""
    sc = S.lattice_scene(128, 128, 0.9, (0.0, 0.0), 7, 0.25 if varied else 0.5, 0.5, jitter=0.08, vel_disc=2.0, constraint_r=52.0,
                         name="dense", cell_size=1.0)
    g, o = _pair(sc.gravity, sc, grid_oracle=True)
    g.set_param(A.PARAM_FUSED, fused)
    g.record_contacts(A.RECORD_PAIRS, 1 << 22)
    for step in range(6):
        r = g.step(1 / 60)
        o.step(1 / 60)
        pg, po = g.pairs_drain(), o.pairs_drain()
        for x, y in zip(pg, po):
            assert np.array_equal(x, y)
        _compare_step(g, o)
    assert r["collisions"] > 8 * 20000
    assert g.kernel_info()["fused_path"] == fused
    cxg, cyg = g.cell_coords()
    cxo, cyo = o.cell_coords()
    assert np.array_equal(cxg, cxo) and np.array_equal(cyg, cyo)


def test_fast_mode_within_tolerance():
    """contact_mode=1 sums a body's contributions in arrival order instead of the reference's pair-loop order: differs only
    by float re-association. Stated tolerance (BASELINE.json north_star): 1e-5 relative over 1 step, on a calm scene (the
    cfg1 transient is chaotic and amplifies ulps within a few substeps)."""
    sc = S.lattice_scene(32, 32, 0.38, (0.0, 0.0), 4, 0.1, 0.2, jitter=0.05, vel_disc=1.0, constraint_r=9.0, name="calm")
    g, o = _pair(sc.gravity, sc)
    g.set_param(A.PARAM_CONTACT_MODE, 1)
    r = g.step(1 / 60)
    o.step(1 / 60)
    assert r["collisions"] > 100
    sg, _ = g.download_bodies()
    so, _ = o.download_bodies()
    for c in ("x", "y"):
        np.testing.assert_allclose(sg["position"][c], so["position"][c], rtol=1e-5, atol=1e-5)


def test_multi_collider_bodies_and_filters():
    """Bodies with 1..4 colliders (offsets, no rotation), sensors, group filters, a static body and a collider-less body."""
    rng = np.random.default_rng(5)
    nb = 600
    b = A.body_descs(nb)
    side = 25
    b["position"]["x"] = (np.arange(nb) % side) * 0.8 + rng.uniform(-0.05, 0.05, nb)
    b["position"]["y"] = (np.arange(nb) // side) * 0.8 + rng.uniform(-0.05, 0.05, nb)
    b["position_old"] = b["position"]
    b["body_type"][::37] = A.BODY_STATIC
    b["gravity_mod"] = rng.uniform(0.5, 1.5, nb).astype(np.float32)
    ncol = rng.integers(0, 5, nb)
    parent = np.repeat(np.arange(nb), ncol)
    nc = len(parent)
    c = A.collider_descs(nc)
    c["radius"] = rng.uniform(0.1, 0.3, nc).astype(np.float32)
    c["offset"]["translation"]["x"] = rng.uniform(-0.3, 0.3, nc).astype(np.float32)
    c["offset"]["translation"]["y"] = rng.uniform(-0.3, 0.3, nc).astype(np.float32)
    c["absolute_transform"]["translation"]["x"] = b["position"]["x"][parent] + c["offset"]["translation"]["x"]
    c["absolute_transform"]["translation"]["y"] = b["position"]["y"][parent] + c["offset"]["translation"]["y"]
    c["is_sensor"][::11] = 1
    c["memberships"][::7] = 0b01
    c["filter"][::5] = 0b10
    c["has_mass_override"][::13] = 1
    c["mass_override"][::13] = 3.5
    sc = S.Scene("multi", gravity=(0.0, -30.0))
    sc.bodies, sc.colliders, sc.col_parent = b, c, parent
    sc.constraints.append((10.0, 10.0, 14.0))
    g, o = _pair(sc.gravity, sc)
    g.record_contacts(A.RECORD_PAIRS, 1 << 20)
    assert g.kernel_info()["n_multi_bodies"] == 0  # topology is built lazily at the first step
    tot = 0
    for _ in range(25):
        g.step(1 / 60)
        o.step(1 / 60)
        for x, y in zip(g.pairs_drain(), o.pairs_drain()):
            assert np.array_equal(x, y)
            tot += len(x)
        _compare_step(g, o)
    assert tot > 500
    assert g.kernel_info()["n_multi_bodies"] > 100


def test_rotating_multi_collider_within_tolerance():
    """Rotation enters the snapshot through sincosf: tolerance-checked, 1e-5 relative after one step."""
    import blobs_b200
    from oracle import oracle_py

    ws = [blobs_b200.World(gravity=(0.0, -30.0)), oracle_py.OracleWorld(gravity=(0.0, -30.0), maintain_spatial_hash=False)]
    for w in ws:
        b = A.body_descs(2)
        b["position"]["x"] = [0.0, 1.3]
        b["position_old"] = b["position"]
        b["rotation"] = [0.3, -1.1]
        bh = w.insert_bodies(b)
        c = A.collider_descs(4)
        c["radius"] = 0.3
        c["offset"]["translation"]["x"] = [0.4, -0.4, 0.4, -0.4]
        par = np.array([0, 0, 1, 1])
        rot = b["rotation"][par]
        # absolute = body.transform() * offset, as a caller would pass it (the first contact pass uses the caller's snapshot)
        c["absolute_transform"]["translation"]["x"] = np.cos(rot) * c["offset"]["translation"]["x"] + b["position"]["x"][par]
        c["absolute_transform"]["translation"]["y"] = np.sin(rot) * c["offset"]["translation"]["x"] + b["position"]["y"][par]
        w.insert_colliders(c, bh[par])
        st = w.body_get(bh[0]).copy()
        st["torque"] = 2.0
        w.body_set(bh[0], st, A.BODY_TORQUE)
    ws[0].step(1 / 60)
    ws[1].step(1 / 60)
    sg, _ = ws[0].download_bodies()
    so, _ = ws[1].download_bodies()
    for f in ("position", "calculated_velocity"):
        for c_ in ("x", "y"):
            np.testing.assert_allclose(sg[f][c_], so[f][c_], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sg["rotation"], so["rotation"], rtol=1e-6)
    np.testing.assert_allclose(sg["angular_velocity"], so["angular_velocity"], rtol=1e-6)
    cg, _ = ws[0].download_colliders()
    co, _ = ws[1].download_colliders()
    for c_ in ("x", "y"):
        np.testing.assert_allclose(cg["desc"]["absolute_transform"]["translation"][c_], co["desc"]["absolute_transform"]["translation"][c_], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("fused", [1, 0])
def test_soft_blobs_springs_and_joints(fused):
    """cfg4 at 64 blobs x 16 bodies: springs (Jacobi) + fixed joints (per-island Gauss-Seidel in slot order, island state in
    shared memory) + contacts. Positions stay bit-exact (anchors are not rotated, so rotation never feeds back into positions);
    rotation is tolerance-checked because of atan2f. fused=1: jointed bodies are advanced inside the joint kernel."""
    sc = S.cfg4(n_blobs=64, k=16, seed=3)
    g, o = _pair(sc.gravity, sc)
    g.set_param(A.PARAM_FUSED, fused)
    g.record_contacts(A.RECORD_PAIRS, 1 << 20)
    for step in range(20):
        g.step(1 / 60)
        o.step(1 / 60)
        for x, y in zip(g.pairs_drain(), o.pairs_drain()):
            assert np.array_equal(x, y)
        _compare_step(g, o)
    info = g.kernel_info()
    assert info["n_islands"] == 64 and info["n_spring_bodies"] == 64 * 16 and info["fused_path"] == fused
    sg, _ = g.download_bodies()
    so, _ = o.download_bodies()
    np.testing.assert_allclose(sg["rotation"], so["rotation"], rtol=1e-5, atol=1e-6)


def test_large_island_and_mixed_bodies():
    """A 130-body chain (island too large for the shared-memory solver -> global-memory fallback, split pipeline), a 12-body
    chain with a static anchor, free spheres and joint_iterations changed at run time, all in one world."""
    import blobs_b200
    from oracle import oracle_py

    ws = [blobs_b200.World(gravity=(0.0, -30.0)), oracle_py.OracleWorld(gravity=(0.0, -30.0), maintain_spatial_hash=False, record_events=False)]
    for w in ws:
        w.constraint_push((0.0, 0.0), 60.0)
        long_chain = [sphere(w, (0.45 * i - 29.0, 5.0 + 0.05 * (i % 3)), r=0.2)[0] for i in range(130)]
        for a, b in zip(long_chain[:-1], long_chain[1:]):
            w.joint_insert(a, b)
        anchor = sphere(w, (0.0, 12.0), r=0.3, body_type=A.BODY_STATIC)[0]
        short = [sphere(w, (0.5 * (i + 1), 12.0), r=0.2)[0] for i in range(12)]
        w.joint_insert(anchor, short[0], (0.1, 0.0), (0.0, 0.0))
        for a, b in zip(short[:-1], short[1:]):
            w.joint_insert(a, b, (0.0, 0.0), (0.0, 0.0), 0.55)
        for i in range(60):
            sphere(w, (-10.0 + 0.35 * i, 8.0 + 0.3 * (i % 4)), r=0.15, velocity_request=(0.5, -3.0))
        w.step(1 / 60, n=4)
        w.set_param(A.PARAM_JOINT_ITERATIONS, 2)
        w.step(1 / 60, n=4)
    _compare_step(ws[0], ws[1])
    sg, _ = ws[0].download_bodies()
    so, _ = ws[1].download_bodies()
    np.testing.assert_allclose(sg["rotation"], so["rotation"], rtol=1e-5, atol=1e-6)
    info = ws[0].kernel_info()
    assert info["n_islands"] == 2 and info["fused_path"] == 0


def test_removal_and_reinsert_mid_simulation():
    """Arena semantics under churn: remove bodies/colliders mid-run, re-insert into recycled slots (generation bump)."""
    sc = S.cfg1(2)
    g, o = _pair(sc.gravity, sc)
    ws = (g, o)
    for w in ws:
        w.step(1 / 60, n=3)
    _, hb = g.download_bodies()
    _, hc = g.download_colliders()
    for w in ws:
        for s in (5, 17, 300, 1023):
            w.remove_body(hb[s])
        for s in (40, 41):
            w.remove_collider(hc[s])  # removes the orphaned parents as well
        w.step(1 / 60, n=2)
    _compare_step(g, o)
    new = []
    for w in ws:
        new.append([sphere(w, (0.1 * i, 5.0 + 0.3 * i), r=0.15, velocity_request=(1.0, -2.0)) for i in range(4)])
        w.step(1 / 60, n=5)
    assert new[0] == new[1]
    assert new[0][0][0] >> 32 == 2
    _compare_step(g, o)
    assert g.body_count() == o.body_count() == 1024 - 6 + 4


def test_body_set_and_translate_between_steps():
    sc = S.cfg1(3)
    g, o = _pair(sc.gravity, sc)
    _, hb = g.download_bodies()
    for w in (g, o):
        w.step(1 / 60)
        st = w.body_get(hb[10]).copy()
        st["position"]["x"] += np.float32(0.25)
        st["velocity_request"]["x"], st["velocity_request"]["y"] = 2.0, 1.0
        st["has_velocity_request"] = 1
        w.body_set(hb[10], st, A.BODY_POSITION | A.BODY_VELOCITY_REQUEST)
        w.body_translate(hb[11], (0.125, -0.5))
        w.body_apply_force(hb[12], (3.0, 9.0))
        st = w.body_get(hb[13]).copy()
        st["body_type"] = A.BODY_STATIC
        w.body_set(hb[13], st, A.BODY_TYPE)
        w.apply_forces(np.tile(np.array([[0.5, 0.25]], dtype=np.float32), (1024, 1)))
        w.step(1 / 60, n=2)
    _compare_step(g, o)


def test_pipelined_host_io_matches_synchronous_calls():
    """blobs_forces_upload_async / blobs_apply_forces_uploaded / blobs_read_body_positions_async / blobs_io_sync (copies on their
    own streams, overlapping the neighbouring steps) against the oracle driven with the synchronous apply_forces + read: the
    positions read back every frame and the final state must be bit-identical."""
    import torch

    sc = S.cfg1(2)
    g, o = _pair(sc.gravity, sc)
    nb, frames = 1024, 12
    pin = torch.cuda.is_available()
    forces = [torch.from_numpy(np.ascontiguousarray(S.uniform(77 + i, 0, 2 * nb).reshape(nb, 2) * np.float32(8.0) - np.float32(4.0))) for i in range(frames + 1)]
    outs = [torch.zeros((nb, 2), dtype=torch.float32) for _ in range(2)]
    if pin:
        forces = [f.pin_memory() for f in forces]
        outs = [x.pin_memory() for x in outs]
    want = []
    for i in range(frames):
        o.apply_forces(forces[i].numpy())
        o.step(1 / 60)
        want.append(o.read_positions().copy())
    got = []
    g.forces_upload_async_ptr(forces[0].data_ptr(), nb)
    with pytest.raises(RuntimeError, match="has not been applied"):
        g.forces_upload_async_ptr(forces[1].data_ptr(), nb)       # at most one pending batch
    for i in range(frames):
        g.apply_forces_uploaded()
        g.forces_upload_async_ptr(forces[i + 1].data_ptr(), nb)   # next frame's input travels under this frame's kernels
        g.step(1 / 60)
        g.io_sync()
        if i:
            got.append(outs[(i - 1) & 1].numpy().copy())          # frame i-1 has arrived
        g.read_positions_async_ptr(outs[i & 1].data_ptr(), nb)    # travels under the next frame's kernels
    g.io_sync()
    got.append(outs[(frames - 1) & 1].numpy().copy())
    for i in range(frames):
        assert np.array_equal(bits(got[i]), bits(want[i])), f"frame {i}"
    g.apply_forces_uploaded()                                     # the last uploaded batch is still applicable
    o.apply_forces(forces[frames].numpy())
    with pytest.raises(RuntimeError, match="no uploaded batch"):
        g.apply_forces_uploaded()
    for w in (g, o):
        w.step(1 / 60)
    _compare_step(g, o)


def test_pipelined_indexed_host_io_matches_synchronous_calls():
    """The distributed (strip) forms of the pipelined host I/O - blobs_forces_indexed_upload_async /
    blobs_apply_forces_indexed_uploaded / blobs_read_owned_positions_async - on an undivided world (every body owned): the
    (slot, position) lists read back every frame and the final state must be bit-identical to the oracle driven with the
    synchronous per-slot calls. Forces are given for a shuffled subset of the slots, as a rank would for the bodies it owns."""
    import torch

    sc = S.cfg1(3)
    g, o = _pair(sc.gravity, sc)
    nb, frames = 1024, 10
    pin = torch.cuda.is_available()
    rng = np.random.default_rng(5)
    slot_lists, force_lists, dense = [], [], []
    for i in range(frames + 1):
        sl = rng.permutation(nb)[: 700 + 20 * i].astype(np.uint32)
        f = np.ascontiguousarray(S.uniform(177 + i, 0, 2 * len(sl)).reshape(len(sl), 2) * np.float32(8.0) - np.float32(4.0))
        d = np.zeros((nb, 2), dtype=np.float32)
        d[sl] = f
        slot_lists.append(torch.from_numpy(sl.view(np.int32).copy()))
        force_lists.append(torch.from_numpy(f))
        dense.append(d)
    cap = nb + 64
    out_slots = [torch.zeros(cap, dtype=torch.int32) for _ in range(2)]
    out_xy = [torch.zeros((cap, 2), dtype=torch.float32) for _ in range(2)]
    out_n = [torch.zeros(1, dtype=torch.int32) for _ in range(2)]
    if pin:
        slot_lists = [x.pin_memory() for x in slot_lists]
        force_lists = [x.pin_memory() for x in force_lists]
        out_slots, out_xy, out_n = ([x.pin_memory() for x in v] for v in (out_slots, out_xy, out_n))
    want = []
    for i in range(frames):
        o.apply_forces(dense[i])          # zero force on the other slots: acc += 0/m leaves every bit as it was (acc is never -0 here)
        o.step(1 / 60)
        want.append(o.read_positions().copy())

    def up(i):
        g.forces_indexed_upload_async_ptr(slot_lists[i].data_ptr(), force_lists[i].data_ptr(), len(slot_lists[i]))

    def landed(k):
        n = int(out_n[k][0])
        assert n == nb
        xy = np.zeros((nb, 2), dtype=np.float32)
        sl = out_slots[k].numpy()[:n]
        assert len(np.unique(sl)) == nb
        xy[sl] = out_xy[k].numpy()[:n]
        return xy

    got = []
    up(0)
    with pytest.raises(RuntimeError, match="has not been applied"):
        up(1)
    with pytest.raises(RuntimeError, match="no uploaded batch"):
        g.apply_forces_uploaded()                                  # an indexed batch is not a per-slot batch
    for i in range(frames):
        g.apply_forces_indexed_uploaded()
        up(i + 1)
        g.step(1 / 60)
        g.io_sync()
        if i:
            got.append(landed((i - 1) & 1))
        g.read_owned_positions_async_ptr(out_slots[i & 1].data_ptr(), out_xy[i & 1].data_ptr(), out_n[i & 1].data_ptr(), cap)
    g.io_sync()
    got.append(landed((frames - 1) & 1))
    for i in range(frames):
        assert np.array_equal(bits(got[i]), bits(want[i])), f"frame {i}"
    g.apply_forces_indexed_uploaded()
    o.apply_forces(dense[frames])
    with pytest.raises(RuntimeError, match="no uploaded indexed batch"):
        g.apply_forces_indexed_uploaded()
    for w in (g, o):
        w.step(1 / 60)
    _compare_step(g, o)


def test_far_outlier_aliases_harmlessly():
    """A body far outside the table's extent wraps around the toroidal grid: still exact."""
    sc = S.cfg1(1)
    g, o = _pair(sc.gravity, sc)
    for w in (g, o):
        sphere(w, (1.0e6, -3.0e5), r=0.2, gravity_mod=0.0)
        sphere(w, (1.0e6 + 0.3, -3.0e5), r=0.2, gravity_mod=0.0)
        w.step(1 / 60, n=3)
    _compare_step(g, o)


@pytest.mark.parametrize("first", ["static", "kinematic", "dynamic"])
def test_first_non_static_body_sees_the_old_dt_ratio(first):
    """physics.rs:327-339: `displacement * (dt / old_dt)` - old_dt is overwritten inside the body loop, so only the FIRST NON-STATIC
    body in arena order sees a ratio != 1, and a static body in slot 0 neither takes that role nor updates old_dt; kinematic bodies
    count as non-static. Bodies carry velocity requests and the delta varies between steps, so a misplaced ratio shows at once."""
    import blobs_b200
    from oracle import oracle_py

    btype = {"static": 1, "kinematic": 2, "dynamic": 0}[first]
    ws = [blobs_b200.World(gravity=(0.0, -30.0)), oracle_py.OracleWorld(gravity=(0.0, -30.0), maintain_spatial_hash=False, record_events=False)]
    for w in ws:
        sphere(w, (0.0, 0.0), r=0.3, body_type=btype, velocity_request=(0.5, 0.25))
        for i in range(6):
            sphere(w, (1.0 + 0.9 * i, 0.1 * i), r=0.3, velocity_request=(3.0 - i, 1.0 + 0.5 * i))
        sphere(w, (-2.0, 0.0), r=0.3, body_type=3, velocity_request=(1.0, 0.0))
        sphere(w, (-3.0, 1.0), r=0.3, body_type=1)
    for delta, sub in ((1 / 60, 8), (1 / 30, 8), (1 / 50, 3), (1 / 60, 8)):
        for w in ws:
            w.set_param(A.PARAM_SUBSTEPS, sub)
            w.step(delta)
        _compare_step(ws[0], ws[1])
        assert np.float32(ws[0].get_param(A.PARAM_OLD_DT)) == np.float32(ws[1].get_param(A.PARAM_OLD_DT))


def test_statics_only_world_keeps_old_dt():
    """a world without any non-static body never overwrites old_dt (physics.rs:338-339 sits behind the is_static `continue`)"""
    import blobs_b200
    from oracle import oracle_py

    ws = [blobs_b200.World(gravity=(0.0, -30.0)), oracle_py.OracleWorld(gravity=(0.0, -30.0), maintain_spatial_hash=False, record_events=False)]
    for w in ws:
        sphere(w, (0.0, 0.0), r=0.3, body_type=1)
        sphere(w, (0.4, 0.0), r=0.3, body_type=1)
        w.step(1 / 60)
        w.step(1 / 30)
    _compare_step(ws[0], ws[1])
    assert np.float32(ws[0].get_param(A.PARAM_OLD_DT)) == np.float32(ws[1].get_param(A.PARAM_OLD_DT)) == np.float32(1.0)


def test_collisions_disabled_and_variable_delta():
    """collisions_enabled=false (physics.rs:25-27) and a delta that changes between steps (Q2 applies to the first body only)."""
    sc = S.cfg1(1)
    g, o = _pair(sc.gravity, sc)
    for w in (g, o):
        w.set_param(A.PARAM_COLLISIONS_ENABLED, 0)
        w.step(1 / 60)
        w.step(1 / 30)
        w.set_param(A.PARAM_COLLISIONS_ENABLED, 1)
        w.set_param(A.PARAM_SUBSTEPS, 3)
        w.step(1 / 50)
    _compare_step(g, o)
    assert np.float32(g.get_param(A.PARAM_OLD_DT)) == np.float32(o.get_param(A.PARAM_OLD_DT))


def test_events_match_reference_channel():
    """CollisionEvent stream (physics.rs:304-311): same (a, b) handles and pre-update velocities, as a multiset per step."""
    sc = S.cfg1(1, n_side=16)
    g, o = _pair(sc.gravity, sc)
    g.record_contacts(A.RECORD_EVENTS, 1 << 18)
    o2 = None
    from oracle import oracle_py

    o2 = oracle_py.OracleWorld(gravity=sc.gravity, maintain_spatial_hash=False, record_events=True)
    S.build(o2, sc)
    n = 0
    for _ in range(30):
        g.step(1 / 60)
        o2.step(1 / 60)
        eg, eo = g.events_drain(), o2.events_drain()
        assert len(eg) == len(eo)
        key = lambda e: np.lexsort((e["col_handle_b"], e["col_handle_a"]))
        eg, eo = eg[key(eg)], eo[key(eo)]
        assert eg.tobytes() == eo.tobytes()
        n += len(eg)
    assert n > 50


def test_events_partial_drains_lose_nothing():
    """blobs_events_drain with a buffer smaller than what was recorded hands the events out in pieces (the Rust shim drains 65 536
    at a time): the concatenation equals one big drain."""
    import ctypes as C

    import blobs_b200

    sc = S.cfg1(1, n_side=16)
    ws = []
    for _ in range(2):
        w = blobs_b200.World(gravity=sc.gravity)
        S.build(w, sc)
        w.record_contacts(A.RECORD_EVENTS, 1 << 16)
        w.step(1 / 60, n=12)
        ws.append(w)
    whole = ws[0].events_drain()
    assert len(whole) > 40
    parts, n = [], C.c_size_t(1 << 30)
    while n.value > 7:
        buf = np.zeros(7, dtype=A.COLLISION_EVENT)
        ws[1]._ck(ws[1]._lib.blobs_events_drain(ws[1]._h, A.ptr(buf), 7, C.byref(n)))
        parts.append(buf[: min(n.value, 7)])
    got = np.concatenate(parts)
    canon = lambda e: np.sort(e.view(np.dtype((np.void, e.dtype.itemsize))))   # recording order depends on which thread got there first
    assert len(got) == len(whole) and np.array_equal(canon(got), canon(whole))
    assert len(ws[1].events_drain()) == 0
    ws[1].step(1 / 60)   # recording restarts cleanly
    ws[0].step(1 / 60)
    assert np.array_equal(canon(ws[1].events_drain()), canon(ws[0].events_drain()))


def test_full_size_cfg2_one_step_vs_grid_oracle():
    """BASELINE config #2 at full size (1 048 576 spheres): one step (8 substeps) vs the cell-list oracle, bit-exact, then
    size-independent properties over more steps: determinism across two GPU runs and bounded penetration."""
    sc = S.cfg2(seed=1)
    g, o = _pair(sc.gravity, sc, grid_oracle=True)
    g.record_contacts(A.RECORD_PAIRS, 1 << 24)
    g.step(1 / 60)
    o.step(1 / 60)
    for x, y in zip(g.pairs_drain(), o.pairs_drain()):
        assert np.array_equal(x, y)
    _compare_step(g, o)
    cxg, cyg = g.cell_coords()
    cxo, cyo = o.cell_coords()
    assert np.array_equal(cxg, cxo) and np.array_equal(cyg, cyo)
    g.record_contacts(A.RECORD_OFF, 0)
    import blobs_b200

    g2 = blobs_b200.World(gravity=sc.gravity)
    S.build(g2, sc)
    g2.step(1 / 60)
    r1 = g.step(1 / 60, n=10)
    r2 = g2.step(1 / 60, n=10)
    assert r1["collisions"] == r2["collisions"] > 0 and r1["nan_detected"] == 0
    p1, p2 = g.read_positions(), g2.read_positions()
    assert p1.tobytes() == p2.tobytes(), "ordered mode must be run-to-run deterministic despite atomic binning"
    assert np.isfinite(p1).all()


def _vs_cpu_checker(scene, checkpoints, params=None):
    """GPU world vs the all-cores CPU checker (oracle/grid_omp.cpp, pinned to the sequential oracle by tests/test_grid_omp.py):
    state bit-equal at every checkpoint, same pair count."""
    import blobs_b200
    from oracle import grid_omp

    g = blobs_b200.World(gravity=scene.gravity)
    S.build(g, scene)
    for k, v in (params or {}).items():
        g.set_param(k, v)
    o = grid_omp.GridOmpWorld(scene)
    done = coll_g = coll_o = 0
    xy = lambda v: np.stack([v["x"], v["y"]], axis=1)
    for cp in checkpoints:
        st = g.step(1 / 60, n=cp - done)
        assert st["nan_detected"] == 0
        coll_g += st["collisions"]
        coll_o += o.step(1 / 60, n=cp - done)["collisions"]
        done = cp
        assert o.coincident == 0, "the CPU checker does not restate the sequential coincident branch"
        sb, _ = g.download_bodies()
        for f, arr in (("position", o.pos), ("position_old", o.pos_old), ("calculated_velocity", o.vel)):
            assert np.array_equal(bits(xy(sb[f])), bits(arr)), f"{f} differs from the CPU checker after {cp} steps"
        assert coll_g == coll_o
    return g, coll_g


def test_full_size_cfg2_twelve_steps_vs_cpu_checker():
    """BASELINE config #2 at full size (1 048 576 spheres), 12 steps = 96 substeps (SURVEY §8d gate: >= 10 steps), library defaults."""
    g, coll = _vs_cpu_checker(S.cfg2(seed=1), (1, 4, 12))
    assert coll > 100_000


@pytest.mark.parametrize("pipeline", ["grid-cooperative-crowded", "lists"])
def test_full_size_contact_rich_start_vs_cpu_checker(pipeline):
    """1 048 576 spheres on a pitch-0.9 lattice (every sphere overlaps its four neighbours from the first substep on, ~2 M pairs per
    substep): the contact-rich paths - cooperative gather, k_crowded hand-over, list rebuilds - checked at scale, 6 steps."""
    g, coll = _vs_cpu_checker(S.cfg2_dense(seed=1, side=1024), (1, 6), params=_PIPELINES[pipeline])
    assert coll > 6 * 8 * 1_000_000


def test_batched_independent_worlds():
    """BASELINE config #3 in small: 24 independent worlds x 256 bodies inside ONE GPU world (world id folded into the cell
    index) vs 24 separate oracle worlds. Every world must behave exactly like its own Physics, including the per-world
    old_dt quirk (Q2) and pair sets (slot offsets are monotone, so the summation order is preserved)."""
    import blobs_b200
    from oracle import oracle_py

    worlds = S.cfg3(n_worlds=24, side=16, seed=11)
    g = blobs_b200.World(gravity=worlds[0].gravity)
    hs = S.build_batch(g, worlds)
    g.record_contacts(A.RECORD_PAIRS, 1 << 20)
    os_ = []
    for sc in worlds:
        o = oracle_py.OracleWorld(gravity=sc.gravity, maintain_spatial_hash=False, record_events=False)
        S.build(o, sc)
        os_.append(o)
    nb = worlds[0].n_bodies
    for step in range(12):
        g.step(1 / 60)
        pg = g.pairs_drain()
        for w, o in enumerate(os_):
            o.step(1 / 60)
            po = o.pairs_drain()
            lo, hi = w * nb, (w + 1) * nb
            for sub in range(8):
                mine = pg[sub][(pg[sub][:, 0] >= lo) & (pg[sub][:, 0] < hi)]
                assert ((mine[:, 1] >= lo) & (mine[:, 1] < hi)).all(), "contact across worlds"
                assert np.array_equal(mine - lo, po[sub]), f"world {w} step {step} substep {sub}"
    sg, _ = g.download_bodies()
    for w, o in enumerate(os_):
        so, _ = o.download_bodies()
        assert_bodies_bit_equal(sg[w * nb:(w + 1) * nb], so)
    assert sum(len(p) for p in pg) > 0


def test_cuda_graph_replay_is_transparent():
    """A whole Physics::integrate call is captured as a CUDA graph and replayed while nothing structural changes. Replays,
    re-captures (dt change, insertion mid-run, profiling on/off) and plain launches must all give the same bits as the oracle."""
    if os.environ.get("BLOBS_TEST_EMU") == "1":
        pytest.skip("the host-compiled test build has no CUDA graphs")
    sc = S.cfg1(2)
    g, o = _pair(sc.gravity, sc)
    import blobs_b200

    plain = blobs_b200.World(gravity=sc.gravity)
    S.build(plain, sc)
    plain.set_param(A.PARAM_GRAPH, 0)
    for w in (g, o, plain):
        w.step(1 / 60, n=6)          # graph: captured on the first step, replayed afterwards (two flavours: last / not last)
        for _ in range(10):          # (the table may be re-dimensioned while the pile collapses: that re-captures too)
            w.step(1 / 60)
    assert g.get_param(A.PARAM_GRAPH_REPLAYS) >= 6 and plain.get_param(A.PARAM_GRAPH_REPLAYS) == 0
    _compare_step(g, o)
    g.profile_enable(True)
    for w in (g, o, plain):
        w.step(1 / 50)               # dt change -> re-capture (and Q2 ratio on the first dynamic body)
        w.step(1 / 50)
        sphere(w, (0.0, 7.0), r=0.1, velocity_request=(0.0, -5.0))   # topology change -> re-capture
        w.step(1 / 50, n=3)
        w.apply_forces(np.tile(np.array([[0.3, 0.1]], dtype=np.float32), (1025, 1)))   # data change only -> replay stays valid
        w.step(1 / 50)
    prof = g.profile_read()
    assert prof["main"][1] >= 8 * 3 and prof["main"][0] > 0.0
    g.profile_enable(False)
    _compare_step(g, o)
    sp, _ = plain.download_bodies()
    so, _ = o.download_bodies()
    assert_bodies_bit_equal(sp, so)


@pytest.mark.parametrize("crowded", [0, 1, 2])
@pytest.mark.parametrize("fused", [1, 0])
@pytest.mark.parametrize("n_small", [30, 200, 1200])
def test_contact_list_overflow_keeps_reference_order(n_small, fused, crowded):
    """More contributions on one body than the 24-entry in-register list: the overflow paths (inline windowed rescan = 0,
    warp-per-body k_crowded = 1, automatic = 2; n_small = 1200 also exceeds k_crowded's 1024-entry sort buffer) must still
    apply them in the reference's pair-loop order (bit-exact), for the big sphere (slot in the middle of the range) and
    for a multi-collider body."""
    import blobs_b200
    from oracle import oracle_py

    ws = [blobs_b200.World(gravity=(0.0, -10.0)), oracle_py.OracleWorld(gravity=(0.0, -10.0), maintain_spatial_hash=False, record_events=False)]
    ws[0].set_param(A.PARAM_CROWDED, crowded)
    ws[0].set_param(A.PARAM_FUSED, fused)
    rng = np.random.default_rng(n_small)
    ang = rng.uniform(0, 2 * np.pi, n_small)
    rad = rng.uniform(0.2, 1.9, n_small)
    for w in ws:
        half = n_small // 2
        for i in range(half):
            sphere(w, (float(rad[i] * np.cos(ang[i])), float(rad[i] * np.sin(ang[i]))), r=0.1)
        sphere(w, (0.0, 0.0), r=2.0)                               # overlaps every small sphere
        for i in range(half, n_small):
            sphere(w, (float(rad[i] * np.cos(ang[i])), float(rad[i] * np.sin(ang[i]))), r=0.1)
        # a two-collider body sitting in the crowd as well
        b = A.body_descs(1)
        b["position"]["x"], b["position"]["y"] = 0.3, 0.2
        b["position_old"] = b["position"]
        bh = w.insert_bodies(b)
        c = A.collider_descs(2)
        c["radius"] = 1.5
        c["offset"]["translation"]["x"] = [0.4, -0.4]
        c["absolute_transform"]["translation"]["x"] = [0.7, -0.1]
        c["absolute_transform"]["translation"]["y"] = [0.2, 0.2]
        w.insert_colliders(c, bh[[0, 0]])
    tot_over = tot_col = 0
    for _ in range(3):
        st = ws[0].step(1 / 60)
        ws[1].step(1 / 60)
        tot_over += st["list_overflow"]
        tot_col += st["collisions"]
        _compare_step(ws[0], ws[1])
    assert tot_over > 0, "the scene must overflow the in-register list"
    assert tot_col == ws[1].step(1 / 60, n=0)["collisions"]   # every pair counted exactly once, whichever kernel resolved it


@pytest.mark.parametrize("crowded", [0, 1, 2])
def test_boundary_shell_vs_grid_oracle(crowded):
    """A lattice block larger than its circle constraint: the clamp projects the corners onto the boundary, which gives a
    dense shell where ~1000 bodies have more than 24 contacts (the regime cfg2 reaches once the block hits the wall)."""
    sc = S.lattice_scene(96, 96, 1.05, (0.0, 0.0), 3, 0.5, 0.5, jitter=0.04, vel_disc=1.0, constraint_r=40.0, name="shell", cell_size=1.0)
    g, o = _pair(sc.gravity, sc, grid_oracle=True)
    g.set_param(A.PARAM_CROWDED, crowded)
    g.record_contacts(A.RECORD_PAIRS, 1 << 22)
    tot_over = tot_col = 0
    for step in range(6):
        st = g.step(1 / 60)
        o.step(1 / 60)
        tot_over += st["list_overflow"]
        tot_col += st["collisions"]
        _compare_step(g, o)
        pg, po = g.pairs_drain(), o.pairs_drain()
        for sub, (x, y) in enumerate(zip(pg, po)):
            assert np.array_equal(x, y), f"pair set differs at step {step} substep {sub}: gpu {len(x)} oracle {len(y)}"
    if int(g.get_param(A.PARAM_POOL)) != 1 and int(g.get_param(A.PARAM_LIST_ACTIVE)) == 0:
        assert tot_over > 500   # per-lane path: these bodies overflow the 24-entry list (the cooperative gather has no such list)
    assert tot_col == o.step(1 / 60, n=0)["collisions"]


@pytest.mark.parametrize("pipeline", list(_PIPELINES))
@pytest.mark.parametrize("scene", ["cfg1", "dense", "shell", "falling"])
def test_every_broadphase_pipeline_gives_the_reference_contact_sets(scene, pipeline):
    """The broadphase strategy (BLOBS_PARAM_LIST: cell grid rebuilt every substep / neighbour lists rebuilt on demand) and the
    contact-resolution variant (per lane / warp-cooperative / k_crowded) never change results: pair set of every substep and the
    state after every step equal the oracle's on a gas (cfg1), an over-full dense pile, a boundary shell and a falling lattice."""
    sc = {"cfg1": lambda: S.cfg1(2),
          "dense": lambda: S.lattice_scene(96, 96, 0.9, (0.0, 0.0), 7, 0.25, 0.5, jitter=0.08, vel_disc=2.0, constraint_r=40.0, name="dense", cell_size=1.0),
          "shell": lambda: S.lattice_scene(96, 96, 1.05, (0.0, 0.0), 3, 0.5, 0.5, jitter=0.04, vel_disc=1.0, constraint_r=40.0, name="shell", cell_size=1.0),
          "falling": lambda: S.lattice_scene(96, 96, 1.05, (0.0, 0.0), 5, 0.3, 0.5, jitter=0.04, vel_disc=1.0, constraint_r=90.0, name="falling", cell_size=1.0)}[scene]()
    g, o = _pair(sc.gravity, sc, grid_oracle=scene != "cfg1")
    for k, v in _PIPELINES[pipeline].items():
        g.set_param(k, v)
    g.record_contacts(A.RECORD_PAIRS, 1 << 22)
    steps = {"cfg1": 12, "dense": 3, "shell": 4, "falling": 12}[scene]
    n = 0
    for step in range(steps):
        st = g.step(1 / 60)
        o.step(1 / 60)
        assert st["nan_detected"] == 0 and st["events_dropped"] == 0
        pg, po = g.pairs_drain(), o.pairs_drain()
        assert len(pg) == len(po) == 8
        for sub, (x, y) in enumerate(zip(pg, po)):
            assert np.array_equal(x, y), f"pair set differs at step {step} substep {sub}: gpu {len(x)} oracle {len(y)}"
            n += len(x)
        _compare_step(g, o)
    assert n > 100
    assert int(g.get_param(A.PARAM_LIST_ACTIVE)) == (1 if pipeline.startswith("lists") else 0)
    if pipeline == "lists" and scene == "falling":   # a lattice falling together keeps its lists for many substeps
        assert g.get_param(A.PARAM_LIST_REBUILDS) < 0.5 * g.get_param(A.PARAM_LIST_SUBSTEPS)
    if pipeline == "lists-tiny-skin":
        assert g.get_param(A.PARAM_LIST_REBUILDS) > 0.5 * g.get_param(A.PARAM_LIST_SUBSTEPS)


def test_debug_data_one_call_snapshot():
    """Physics::debug_data (debug.rs:34-91) through blobs_debug_data: arena order for all four lists (a removed and re-inserted
    spring re-uses its slot), joint / spring endpoints = live body positions, collider transforms = the live snapshot
    (translation bit-exact vs the oracle, matrix = M(rotation) * offset.matrix2 once a substep has run)."""
    import blobs_b200
    from oracle import oracle_py

    g = blobs_b200.World(gravity=(0.0, -5.0))
    o = oracle_py.OracleWorld(gravity=(0.0, -5.0), maintain_spatial_hash=False, record_events=False)
    hs = []
    for w in (g, o):
        b = [sphere(w, (0.0, 0.0), r=0.3)[0], sphere(w, (1.0, 0.2), r=0.3)[0], sphere(w, (2.1, 0.0), r=0.25)[0], sphere(w, (0.5, 1.4), r=0.2)[0]]
        # a body with two colliders, one of them with a rotated offset
        bd = A.body_descs(1)
        bd["position"]["x"], bd["position"]["y"] = 4.0, 1.0
        bd["position_old"] = bd["position"]
        mb = w.insert_bodies(bd)
        cd = A.collider_descs(2)
        cd["radius"] = 0.2
        cd["shape_radius"] = [0.2, 0.35]
        cd["offset"]["translation"]["x"] = [0.3, -0.3]
        cd["offset"]["x_axis"]["x"], cd["offset"]["x_axis"]["y"] = [1.0, 0.0], [0.0, 1.0]      # second offset rotated by 90 degrees
        cd["offset"]["y_axis"]["x"], cd["offset"]["y_axis"]["y"] = [0.0, -1.0], [1.0, 0.0]
        cd["absolute_transform"]["translation"]["x"] = [4.3, 3.7]
        cd["absolute_transform"]["translation"]["y"] = [1.0, 1.0]
        w.insert_colliders(cd, mb[[0, 0]])
        j0 = w.joint_insert(b[0], b[1])                 # the joint makes both bodies rotate (physics.rs:455-469)
        s0 = w.spring_insert(b[1], b[2], 1.0, 50.0, 1.0)
        s1 = w.spring_insert(b[2], b[3], 1.5, 20.0, 0.5)
        w.spring_remove(s0)
        s2 = w.spring_insert(b[3], b[0], 1.2, 30.0, 0.5)   # LIFO: lands in s0's slot, so it iterates BEFORE s1
        hs.append((b, int(mb[0]), j0, s1, s2))
    assert hs[0] == hs[1]
    b, mb, _, _, _ = hs[0]
    before = g.debug_data()
    assert np.allclose(before["colliders"][4:, :4], [[1, 0, 0, 1], [1, 0, 0, 1]])    # caller's absolute_transform until a substep runs
    for _ in range(5):
        g.step(1 / 60)
        o.step(1 / 60)
    d = g.debug_data()
    sb, hb = g.download_bodies()
    ob, _ = o.download_bodies()
    oc, och = o.download_colliders()
    slot = lambda h: int(h) & 0xFFFFFFFF
    live = hb != 0
    assert d["bodies"].shape == (5, 6) and d["joints"].shape == (1, 4) and d["colliders"].shape == (6, 6) and d["springs"].shape == (2, 4)
    px, py, rot = sb["position"]["x"][live], sb["position"]["y"][live], sb["rotation"][live]
    assert np.array_equal(bits(d["bodies"][:, 4]), bits(px)) and np.array_equal(bits(d["bodies"][:, 5]), bits(py))
    assert np.array_equal(bits(px), bits(ob["position"]["x"][live]))
    assert np.allclose(d["bodies"][:, 0], np.cos(rot), atol=1e-6) and np.allclose(d["bodies"][:, 1], np.sin(rot), atol=1e-6)
    assert np.allclose(d["bodies"][:, 2], -np.sin(rot), atol=1e-6) and np.allclose(d["bodies"][:, 3], np.cos(rot), atol=1e-6)
    assert np.abs(rot[:2]).max() > 0, "the jointed bodies must have rotated"
    pos = lambda h: (sb["position"]["x"][slot(h)], sb["position"]["y"][slot(h)])
    assert np.array_equal(bits(d["joints"][0]), bits(np.array([*pos(b[0]), *pos(b[1])], dtype=np.float32)))
    assert np.array_equal(bits(d["springs"][0]), bits(np.array([*pos(b[3]), *pos(b[0])], dtype=np.float32)))   # re-used slot first
    assert np.array_equal(bits(d["springs"][1]), bits(np.array([*pos(b[2]), *pos(b[3])], dtype=np.float32)))
    olive = och != 0
    ot = oc["desc"]["absolute_transform"][olive]
    assert np.array_equal(bits(d["colliders"][:, 4]), bits(ot["translation"]["x"])) and np.array_equal(bits(d["colliders"][:, 5]), bits(ot["translation"]["y"]))
    want_m = np.stack([ot["x_axis"]["x"], ot["x_axis"]["y"], ot["y_axis"]["x"], ot["y_axis"]["y"]], axis=1)
    assert np.allclose(d["colliders"][:, :4], want_m, atol=1e-6)
    assert np.allclose(d["colliders"][5, :4], [0, 1, -1, 0], atol=1e-6)          # M(0) * the 90-degree offset
    assert np.allclose(d["collider_radius"], [0.3, 0.3, 0.25, 0.2, 0.2, 0.35])


def _brute_force_query(cx, cy, qr, px, py, pr):
    """SpatialHash::query's hit test (spatial.rs:171-177) over ALL points, in f32 op by op."""
    f = np.float32
    dx, dy = (px - f(cx)).astype(f), (py - f(cy)).astype(f)
    d2 = ((dx * dx).astype(f) + (dy * dy).astype(f)).astype(f)
    dist = (f(qr) + pr).astype(f)
    return d2 <= (dist * dist).astype(f)


def test_scene_queries_served_from_the_grid():
    """blobs_query_circles (SpatialHash::query, spatial.rs:155-195 + the stubbed QueryFilter): hits from the GPU
    broadphase table == brute force over every live collider snapshot with the reference's inclusive f32 hit test, for radii
    from 0 to several cells; == the oracle's SpatialHash::query where that one's 3x3 window is complete; every filter."""
    import blobs_b200
    from oracle import oracle_py

    sc = S.cfg1(seed=4)
    g, o = _pair(sc.gravity, sc)
    extra = []
    for w in (g, o):
        hs = [sphere(w, (0.5, 0.5), r=0.15, body_type=A.BODY_STATIC), sphere(w, (-1.0, 2.0), r=0.12, body_type=A.BODY_KINEMATIC_POSITION),
              sphere(w, (1.5, 3.0), r=0.18, col={"is_sensor": 1}), sphere(w, (-2.0, 4.0), r=0.1, col={"memberships": 0b0100, "filter": 0b0010})]
        extra.append(hs)
    assert extra[0] == extra[1]
    (static_b, static_c), (kin_b, kin_c), (sens_b, sens_c), (grp_b, grp_c) = extra[0]
    for _ in range(12):
        g.step(1 / 60)
        o.step(1 / 60)
    oc, och = o.download_colliders()
    live = och != 0
    px, py = oc["desc"]["absolute_transform"]["translation"]["x"][live], oc["desc"]["absolute_transform"]["translation"]["y"][live]
    pr, handles = oc["desc"]["radius"][live], och[live]
    rng = np.random.default_rng(11)
    nq = 300
    centres = np.stack([rng.uniform(-7, 7, nq), rng.uniform(-8, 8, nq)], axis=1).astype(np.float32)
    centres[:20] = np.stack([px[:20], py[:20]], axis=1)          # some queries exactly on a collider
    radii = rng.choice([0.0, 0.05, 0.3, 1.0, 3.0], nq).astype(np.float32)
    radii[5], radii[6] = -1.0, np.nan                            # no hits, no crash
    off, hits = g.query_circles(centres, radii)
    assert off[0] == 0 and off[-1] == len(hits)
    total = 0
    for q in range(nq):
        want = handles[_brute_force_query(centres[q, 0], centres[q, 1], radii[q], px, py, pr)] if radii[q] >= 0 else handles[:0]
        got = hits[off[q]:off[q + 1]]
        assert np.array_equal(got, np.sort(want)), f"query {q}: centre {centres[q]} r {radii[q]}: {len(got)} hits, brute force {len(want)}"
        total += len(got)
    assert total > 2000
    # the reference's own query: identical where its 3x3 window cannot miss anything (query radius + point radius <= cell size)
    sh = oracle_py.OracleSpatialHash(2.0)
    for i in range(len(px)):
        sh.insert_with_id(int(handles[i]), (float(px[i]), float(py[i])), float(pr[i]))
    for q in range(0, nq, 7):
        if not (0 <= radii[q] <= 1.0):
            continue
        ref_ids = sorted(pid for pid, _ in sh.query((float(centres[q, 0]), float(centres[q, 1])), float(radii[q])))
        assert ref_ids == [int(h) for h in hits[off[q]:off[q + 1]]]
    # filters (query_filter.rs:27-108), on one big query that sees everything
    everything = lambda **kw: set(int(h) for h in g.query_circles([[0.0, 0.0]], [100.0], **kw)[1])
    full = everything()
    assert full == set(int(h) for h in handles)
    assert everything(flags=A.QUERY_EXCLUDE_SENSORS) == full - {sens_c}
    assert everything(flags=A.QUERY_EXCLUDE_SOLIDS) == {sens_c}
    assert everything(flags=A.QUERY_EXCLUDE_FIXED) == full - {static_c}
    assert everything(flags=A.QUERY_EXCLUDE_KINEMATIC) == full - {kin_c}
    assert everything(flags=A.QUERY_EXCLUDE_DYNAMIC) == {static_c, kin_c}
    assert everything(flags=A.QUERY_EXCLUDE_DYNAMIC | A.QUERY_EXCLUDE_KINEMATIC) == {static_c}            # ONLY_FIXED
    assert everything(exclude_collider=grp_c) == full - {grp_c}
    assert everything(exclude_rigid_body=kin_b) == full - {kin_c}
    assert everything(groups=(0b0010, 0b0100)) == full                                                     # compatible with the special collider too
    assert everything(groups=(0b0001, 0b0001)) == full - {grp_c}                                           # groups.rs:52-57 fails for it
    # capacity protocol of the C ABI
    import ctypes as C
    offs = np.zeros(2, dtype=np.uint64)
    nh = C.c_size_t(0)
    c1, r1 = np.array([[0.0, 0.0]], dtype=np.float32), np.array([100.0], dtype=np.float32)
    small = np.zeros(4, dtype=np.uint64)
    rc = g._lib.blobs_query_circles(g._h, 1, A.ptr(c1), A.ptr(r1), None, A.ptr(offs), A.ptr(small), 4, C.byref(nh))
    assert rc == A.ERR_CAPACITY and nh.value == len(full) and not small.any()
