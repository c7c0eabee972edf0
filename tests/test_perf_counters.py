"""The reference's process-global perf-counter registry (blobs/src/perf_counters.rs) behind the C ABI. Host-only code, so
the registry semantics are checked on the CPU through the real libblobs_b200.so; that Physics::step feeds "collisions"
(physics.rs:316) needs a stepping world and is checked under -m gpu (and through the host-compiled kernels, -m emu)."""
import pytest


class RefCounters:
    """line-by-line restatement of PerfCounters (perf_counters.rs:17-50) as the checker"""

    def __init__(self):
        self.c = {}

    def update(self, name, count):
        self.c.setdefault(name, [0, 0.0])[0] = count

    def inc(self, name, by):
        cur = self.c[name][0] if name in self.c else 0
        self.update(name, cur + by)

    def new_frame(self, delta):
        for v in self.c.values():
            v[1] = v[1] * (1.0 - delta) + float(v[0]) * delta
            v[0] = 0

    def get(self, name):
        return tuple(self.c[name]) if name in self.c else (0, 0.0)


def test_registry_semantics_match_the_reference(blobs):
    from blobs_b200 import perf_counters as pc

    pc.reset_perf_counters()
    ref = RefCounters()
    assert pc.get_perf_counter("collisions") == (0, 0.0) and pc.counters() == {}
    script = [("inc", "collisions", 7), ("inc", "collisions", 5), ("set", "query", 3), ("frame", 1 / 60), ("inc", "collisions", 100),
              ("frame", 0.25), ("frame", 0.25), ("set", "query", 9), ("inc", "query", 1), ("inc", "fresh", 2), ("frame", 1 / 144)]
    for op in script:
        if op[0] == "inc":
            pc.perf_counter_inc(op[1], op[2]); ref.inc(op[1], op[2])
        elif op[0] == "set":
            pc.perf_counter(op[1], op[2]); ref.update(op[1], op[2])
        else:
            pc.perf_counters_new_frame(op[1]); ref.new_frame(op[1])
        for name in ("collisions", "query", "fresh", "absent"):
            assert pc.get_perf_counter(name) == ref.get(name), (op, name)   # f64 arithmetic in the same order: exact
    assert pc.counters() == {k: tuple(v) for k, v in ref.c.items()}
    assert list(pc.counters()) == sorted(ref.c)
    pc.reset_perf_counters()
    assert pc.counters() == {} and pc.get_perf_counter("collisions") == (0, 0.0)


def test_counter_listing_reports_short_buffers(blobs):
    import ctypes as C

    from blobs_b200 import _abi as A
    from blobs_b200 import perf_counters as pc
    from blobs_b200._lib import load

    pc.reset_perf_counters()
    pc.perf_counter("a-rather-long-counter-name", 1)
    lib = load()
    small = C.create_string_buffer(4)
    assert lib.blobs_perf_counter_at(0, small, len(small), None, None) == A.ERR_CAPACITY
    assert lib.blobs_perf_counter_at(1, small, len(small), None, None) == A.ERR_INVALID
    assert lib.blobs_perf_counter_at(0, None, 0, None, None) == A.OK
    pc.reset_perf_counters()


def check_step_feeds_collisions():
    """two overlapping balls: every substep of every step counts one pair; 'collisions' accumulates them until new_frame"""
    from blobs_b200 import perf_counters as pc
    from blobs_b200.physics import Affine2, ColliderBuilder, Physics, RigidBodyBuilder

    pc.reset_perf_counters()
    physics = Physics(gravity=(0.0, 0.0))
    for x in (0.0, 0.05):
        rbd = physics.insert_rbd(RigidBodyBuilder().position((x, 0.0)).build())
        physics.insert_collider_with_parent(ColliderBuilder().radius(0.5).absolute_transform(Affine2.from_translation((x, 0.0))).build(), rbd)
    total = 0
    for _ in range(3):
        total += physics.step(1 / 60)["collisions"]
    assert total > 0
    assert pc.get_perf_counter("collisions") == (total, 0.0)
    pc.perf_counters_new_frame(1 / 60)
    assert pc.get_perf_counter("collisions") == (0, total * (1 / 60))
    physics.collisions_enabled = False
    physics.step(1 / 60)
    assert pc.get_perf_counter("collisions")[0] == 0
    pc.reset_perf_counters()


@pytest.mark.gpu
def test_step_feeds_the_collisions_counter():
    check_step_feeds_collisions()
