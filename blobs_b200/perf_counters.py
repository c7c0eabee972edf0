"""Mirror of the reference's `blobs::perf_counters` module (blobs/src/perf_counters.rs:52-87) over the C ABI: one
process-global registry of named counters, fed by `Physics::step` ("collisions", physics.rs:316) and read by the
application's perf panel (demo/src/main.rs:291-300). Same function names as the reference."""
import ctypes as C

from ._lib import load


def _name(n):
    return n.encode() if isinstance(n, str) else bytes(n)


def perf_counter(counter_name, count):  # perf_counters.rs:66-69
    load().blobs_perf_counter(_name(counter_name), int(count))


def perf_counter_inc(counter_name, inc):  # perf_counters.rs:71-76
    load().blobs_perf_counter_inc(_name(counter_name), int(inc))


def perf_counters_new_frame(delta):  # perf_counters.rs:56-59
    load().blobs_perf_counters_new_frame(float(delta))


def reset_perf_counters():  # perf_counters.rs:61-64
    load().blobs_perf_counters_reset()


def get_perf_counter(counter_name):  # perf_counters.rs:78-81 -> (count, decayed_average); (0, 0.0) when absent
    c, a = C.c_uint64(), C.c_double()
    load().blobs_perf_counter_get(_name(counter_name), C.byref(c), C.byref(a))
    return c.value, a.value


def counters():
    """PerfCounters::global().counters as {name: (count, decayed_average)} (perf_counters.rs:6-9)."""
    lib = load()
    out = {}
    buf = C.create_string_buffer(256)
    for i in range(lib.blobs_perf_counter_count()):
        c, a = C.c_uint64(), C.c_double()
        if lib.blobs_perf_counter_at(i, buf, len(buf), C.byref(c), C.byref(a)) == 0:
            out[buf.value.decode()] = (c.value, a.value)
    return out
