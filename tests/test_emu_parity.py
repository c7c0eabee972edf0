"""Kernel LOGIC on the CPU: the parity tests of test_gpu_parity.py / test_physics_api.py, run against the host-compiled
build of the very same CUDA sources (tests/emu: one fiber per CUDA thread, warp collectives and shared memory emulated)
and compared bit-for-bit with the oracle. This is what lets a kernel change be checked in a container without a GPU; it
says nothing about speed and is not a product path (blobs_b200 never loads the emulation library - see emu_loader.py).
The `-m gpu` runs of the same tests on the B200 remain the parity gate proper.

Only cases that finish in a few seconds are listed; `BLOBS_TEST_EMU=1 python -m pytest tests -m gpu -k ...` runs any GPU
test this way (the whole parity suite takes ~8 minutes)."""
import pytest

from . import test_gpu_parity as T
from . import test_perf_counters as PC
from . import test_physics_api as P
from . import test_rigid_body_api as RB
from .emu_loader import emulated

pytestmark = pytest.mark.emu

# knob sets: library defaults (automatic kernel selection) / every optional kernel path forced on for every warp
DEFAULT = {}
FORCED = {"BLOBS_B200_POOL": "1", "BLOBS_B200_POOL_MIN": "1", "BLOBS_B200_CROWDED": "1"}


def _skip_if_redundant(case, knobs):
    """The 30 s combinations whose paths the other cases already walk run only with BLOBS_TEST_SLOW=1 (the CPU suite is meant
    to finish in a few minutes); the default-knob twin of each still runs."""
    import os

    if case[0] == "batched-worlds" and knobs is FORCED and os.environ.get("BLOBS_TEST_SLOW") != "1":
        pytest.skip("redundant combination; set BLOBS_TEST_SLOW=1")

CASES = [
    ("overflow30-fused-inline", T.test_contact_list_overflow_keeps_reference_order, dict(n_small=30, fused=1, crowded=0)),
    ("overflow30-split-crowded", T.test_contact_list_overflow_keeps_reference_order, dict(n_small=30, fused=0, crowded=1)),
    ("overflow200-fused-crowded", T.test_contact_list_overflow_keeps_reference_order, dict(n_small=200, fused=1, crowded=1)),
    ("overflow200-split-auto", T.test_contact_list_overflow_keeps_reference_order, dict(n_small=200, fused=0, crowded=2)),
    ("overflow1200-fused-crowded", T.test_contact_list_overflow_keeps_reference_order, dict(n_small=1200, fused=1, crowded=1)),
    ("overflow1200-fused-inline", T.test_contact_list_overflow_keeps_reference_order, dict(n_small=1200, fused=1, crowded=0)),
    ("multi-collider", T.test_multi_collider_bodies_and_filters, {}),
    ("rotating-multi-collider", T.test_rotating_multi_collider_within_tolerance, {}),
    ("removal-reinsert", T.test_removal_and_reinsert_mid_simulation, {}),
    ("body-set-translate", T.test_body_set_and_translate_between_steps, {}),
    ("far-outlier", T.test_far_outlier_aliases_harmlessly, {}),
    ("pipelined-host-io", T.test_pipelined_host_io_matches_synchronous_calls, {}),
    ("pipelined-indexed-host-io", T.test_pipelined_indexed_host_io_matches_synchronous_calls, {}),
    ("collisions-disabled-variable-delta", T.test_collisions_disabled_and_variable_delta, {}),
    ("first-non-static-static-in-slot-0", T.test_first_non_static_body_sees_the_old_dt_ratio, dict(first="static")),
    ("first-non-static-kinematic-in-slot-0", T.test_first_non_static_body_sees_the_old_dt_ratio, dict(first="kinematic")),
    ("statics-only-keeps-old-dt", T.test_statics_only_world_keeps_old_dt, {}),
    ("events", T.test_events_match_reference_channel, {}),
    ("large-island", T.test_large_island_and_mixed_bodies, {}),
    ("fast-mode", T.test_fast_mode_within_tolerance, {}),
    ("batched-worlds", T.test_batched_independent_worlds, {}),
    ("soft-blobs-fused", T.test_soft_blobs_springs_and_joints, dict(fused=1)),
    ("debug-data", T.test_debug_data_one_call_snapshot, {}),
    ("scene-queries", T.test_scene_queries_served_from_the_grid, {}),
    ("physics-api-balls", P.test_balls_demo_flow, {}),
    ("physics-api-joints-springs-panics", P.test_joints_springs_and_panics, {}),
    ("perf-counter-collisions", PC.check_step_feeds_collisions, {}),
    ("removal-semantics-event-ring", RB.check_removal_semantics_and_event_ring, {}),
]


@pytest.mark.parametrize("knobs", [DEFAULT, FORCED], ids=["default", "forced-pool-crowded"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_kernel_logic_on_cpu(case, knobs, monkeypatch):
    _, fn, kw = case
    _skip_if_redundant(case, knobs)
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)   # read by World's constructor (same meaning as the BLOBS_PARAM_* knobs)
    with emulated():
        fn(**kw)


def test_results_do_not_depend_on_the_schedule():
    """Same cases with the emulator visiting CTAs, warps and lanes in a seeded RANDOM order (BLOBS_EMU_SEED, read when the
    library is loaded, hence the subprocess): atomics then hand out different ranks and cells hold their records in another
    order, yet every result must stay bit-identical to the oracle's."""
    import os
    import subprocess
    import sys

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, BLOBS_EMU_SEED="20261017")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(repo, "tests", "test_emu_parity.py"), "-q", "-x", "-p", "no:cacheprovider",
                        "-k", "(test_kernel_logic_on_cpu and (overflow200 or overflow1200-fused-crowded or multi-collider or large-island or events or removal))"],
                       capture_output=True, text=True, timeout=900, env=env, cwd=repo)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and " passed" in r.stdout
