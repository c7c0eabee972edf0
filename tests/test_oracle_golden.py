"""Pins the CPU oracle against every vector the reference's own tests hold for this path
(SURVEY.md §4, §8c): the five SpatialHash tests (blobs/src/tests.rs:26-90) and the three transform
tests (blobs/src/collider.rs:340-389). The reference has NO test that calls Physics::step, so the step
itself is 'parity unpinned' by reference tests; tests/test_oracle_hand.py adds hand-derived vectors."""
import math

import numpy as np


def _create_spatial_hash(o):  # tests.rs:16-24
    sh = o.OracleSpatialHash(100.0)
    for p in [(50.0, 50.0), (150.0, 150.0), (250.0, 250.0), (350.0, 350.0), (450.0, 450.0)]:
        sh.insert(p, 1.0)
    return sh


def test_get_cell_coords(oracle):  # tests.rs:26-34
    sh = _create_spatial_hash(oracle)
    assert sh.get_cell_coords((50.0, 50.0)) == (0, 0)
    assert sh.get_cell_coords((150.0, 150.0)) == (1, 1)
    assert sh.get_cell_coords((250.0, 250.0)) == (2, 2)
    assert sh.get_cell_coords((350.0, 350.0)) == (3, 3)
    assert sh.get_cell_coords((450.0, 450.0)) == (4, 4)


def test_insert(oracle):  # tests.rs:36-53
    sh = oracle.OracleSpatialHash(100.0)
    pid = sh.insert((50.5, -25.5), 1.0)
    assert pid == 0
    assert sh.next_id == 1
    cell = sh.get_cell_coords((50.5, -25.5))
    assert cell == (0, -1)  # floor of a negative coordinate
    assert sh.cell_population(cell) > 0
    pt = sh.point(pid)
    assert abs(pt[0] - 50.5) < 1e-6 and abs(pt[1] + 25.5) < 1e-6


def test_insert_and_query(oracle):  # tests.rs:55-65
    sh = oracle.OracleSpatialHash(1.0)
    p1 = sh.insert((0.0, 0.0), 0.0)
    p2 = sh.insert((2.0, 2.0), 0.0)
    res = sh.query((1.0, 1.0), 1.5)
    assert len(res) == 2
    assert {r[0] for r in res} == {p1, p2}


def test_move_point(oracle):  # tests.rs:67-79
    sh = oracle.OracleSpatialHash(1.0)
    p = sh.insert((0.0, 0.0), 0.0)
    assert sh.move_point(p, (2.0, 2.0))
    res = sh.query((1.0, 1.0), 1.5)
    assert len(res) == 1
    assert res[0][0] == p and res[0][1][0] == 2.0 and res[0][1][1] == 2.0


def test_remove(oracle):  # tests.rs:81-90
    sh = oracle.OracleSpatialHash(1.0)
    p = sh.insert((0.0, 0.0), 0.0)
    sh.remove(p)
    assert sh.query((0.0, 0.0), 1.5) == []


def _aff(a):
    return np.array([a.x_axis.x, a.x_axis.y, a.y_axis.x, a.y_axis.y, a.translation.x, a.translation.y], dtype=np.float32)


def test_simple_body_transform(oracle):  # collider.rs:340-351
    t = _aff(oracle.body_transform(0.0, (2.0, 3.0)))
    np.testing.assert_array_equal(t, np.array([1, 0, 0, 1, 2, 3], dtype=np.float32))


def test_body_transform(oracle):  # collider.rs:353-365 — matrix2 == Mat2::from_angle(PI/2)
    t = _aff(oracle.body_transform(np.float32(math.pi / 2), (2.0, 3.0)))
    s, c = np.sin(np.float32(math.pi / 2)), np.cos(np.float32(math.pi / 2))
    np.testing.assert_allclose(t[:4], [c, s, -s, c], rtol=0, atol=1e-7)


def test_collider_offset(oracle):  # collider.rs:367-389
    A = oracle.A
    off = A.Affine2(A.Vec2(1.0, 0.0), A.Vec2(0.0, 1.0), A.Vec2(1.0, 1.0))
    body = oracle.body_transform(np.float32(math.pi / 2), (2.0, 3.0))
    got = _aff(oracle.affine_mul(body, off))
    # Mat2::from_angle(pi/2) * (1,1) + (2,3) = (cos - sin, sin + cos) + (2,3) ~= (1, 4)
    np.testing.assert_allclose(got[4:], [1.0, 4.0], atol=1e-6)
    np.testing.assert_allclose(got[:4], _aff(body)[:4], atol=0)
