"""CPU-only: the checkers check each other.

1. The sequential GRID oracle (Physics::grid_collisions: cell-list candidates resolved by the very same resolve_pair(), in the
   brute-force loop's order) equals the BRUTE-FORCE oracle (physics.rs:241-317 restated) bit for bit — every 1M-scale and
   dense-scene GPU parity test hangs on the grid oracle, so it is pinned to the reference algorithm here.
2. The all-cores cell-list restatement (oracle/grid_omp.cpp: gather per body, contributions in ascending partner slot) equals
   the sequential oracle bit for bit; it is what makes full-size multi-step parity checks and the labelled CPU number possible."""
import numpy as np
import pytest

from blobs_b200 import scenes as S

from .helpers import bits


def _oracle(scene, grid):
    from oracle import oracle_py

    o = oracle_py.OracleWorld(gravity=scene.gravity, grid_pairs=grid, maintain_spatial_hash=False, record_events=False)
    S.build(o, scene)
    return o


def _scenes():
    return {
        "cfg1": S.cfg1(1),
        "cfg1-seed2-small": S.cfg1(2, n_side=16),
        "falling-lattice": S.lattice_scene(48, 48, 1.05, (0.0, 0.0), 1, 0.5, 0.5, jitter=0.04, vel_disc=1.0, constraint_r=40.0, name="fall", cell_size=1.0),
        "dense-pile-4k": S.lattice_scene(64, 64, 0.9, (0.0, 0.0), 7, 0.25, 0.5, jitter=0.08, vel_disc=2.0, constraint_r=26.0, name="dense", cell_size=1.0),
    }


@pytest.mark.parametrize("name", ["cfg1", "dense-pile-4k"])
def test_grid_oracle_equals_brute_force_oracle(name):
    sc = _scenes()[name]
    a, b = _oracle(sc, False), _oracle(sc, True)
    steps = 25 if name == "cfg1" else 6
    for _ in range(steps):
        a.step(1 / 60)
        b.step(1 / 60)
        pa, pb = a.pairs_drain(), b.pairs_drain()
        assert len(pa) == len(pb) == 8
        for x, y in zip(pa, pb):
            assert np.array_equal(x, y), "pair set per substep"
    sa, _ = a.download_bodies()
    sb, _ = b.download_bodies()
    for f in ("position", "position_old", "calculated_velocity"):
        for c in ("x", "y"):
            assert np.array_equal(bits(sa[f][c]), bits(sb[f][c])), f
    assert a.step(1 / 60, n=0)["collisions"] == b.step(1 / 60, n=0)["collisions"] > 1000
    assert a.coincident_total() == b.coincident_total() == 0


@pytest.mark.parametrize("name", ["cfg1", "cfg1-seed2-small", "falling-lattice", "dense-pile-4k"])
@pytest.mark.parametrize("threads", [1, 0], ids=["1-thread", "all-threads"])
def test_all_cores_restatement_equals_sequential_oracle(name, threads):
    from oracle import grid_omp

    sc = _scenes()[name]
    o = _oracle(sc, name != "cfg1")
    g = grid_omp.GridOmpWorld(sc, threads=threads)
    steps = {"cfg1": 25, "cfg1-seed2-small": 40, "falling-lattice": 30, "dense-pile-4k": 8}[name]
    for chunk in (1, steps - 1):   # one call of 1 step, one call of many: old_dt carries over
        o.step(1 / 60, n=chunk)
        g.step(1 / 60, n=chunk)
        so, _ = o.download_bodies()
        for f, arr in (("position", g.pos), ("position_old", g.pos_old), ("calculated_velocity", g.vel), ("acceleration", g.acc)):
            for k, c in enumerate(("x", "y")):
                assert np.array_equal(bits(so[f][c]), bits(arr[:, k])), f"{f}.{c}"
        co, _ = o.download_colliders()
        for k, c in enumerate(("x", "y")):
            assert np.array_equal(bits(co["desc"]["absolute_transform"]["translation"][c]), bits(g.snap[:, k])), "snapshot"
    assert g.collisions == o.step(1 / 60, n=0)["collisions"] > 100
    assert g.coincident == 0
