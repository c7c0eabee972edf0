"""GPU parity of the opt-in kernel variants written after round 1's GPU minutes were spent (k_tile: BLOBS_PARAM_TUNE 11, per-lane
and warp-pooled): the scene-level parity tests of test_gpu_parity.py again, through those variants, against the same oracle.
Collected LAST on purpose (file name): these variants had only run on the host-compiled build (tests/emu) when they were
committed, and `pytest -x` must reach every test of the default path before it reaches them."""
import pytest

from . import test_gpu_parity as T

pytestmark = pytest.mark.gpu

FORCED = {"BLOBS_B200_POOL": "1", "BLOBS_B200_POOL_MIN": "1", "BLOBS_B200_CROWDED": "1"}

CASES = [
    ("cfg1-every-substep", T.test_cfg1_pairs_and_positions_every_substep, dict(seed=1)),
    ("dense-pile-fused", T.test_dense_pile_vs_grid_oracle, dict(fused=1, varied=False)),
    ("dense-pile-varied", T.test_dense_pile_vs_grid_oracle, dict(fused=1, varied=True)),
    ("multi-collider", T.test_multi_collider_bodies_and_filters, {}),
    ("soft-blobs-fused", T.test_soft_blobs_springs_and_joints, dict(fused=1)),
    ("removal-reinsert", T.test_removal_and_reinsert_mid_simulation, {}),
    ("far-outlier", T.test_far_outlier_aliases_harmlessly, {}),
    ("batched-worlds", T.test_batched_independent_worlds, {}),
    ("cuda-graph-replay", T.test_cuda_graph_replay_is_transparent, {}),
    ("overflow200-crowded", T.test_contact_list_overflow_keeps_reference_order, dict(n_small=200, fused=1, crowded=1)),
    ("overflow1200-inline", T.test_contact_list_overflow_keeps_reference_order, dict(n_small=1200, fused=1, crowded=0)),
    ("boundary-shell-auto", T.test_boundary_shell_vs_grid_oracle, dict(crowded=2)),
    ("full-size-cfg2", T.test_full_size_cfg2_one_step_vs_grid_oracle, {}),
]


@pytest.mark.parametrize("knobs", [{}, FORCED], ids=["default", "forced-pool-crowded"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tile_kernel_on_gpu(case, knobs, monkeypatch):
    name, fn, kw = case
    monkeypatch.setenv("BLOBS_B200_TUNE", "11")   # read by World's constructor, same meaning as BLOBS_PARAM_TUNE
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    fn(**kw)


SMALL_TILE_CASES = [c for c in CASES if c[0] in ("cfg1-every-substep", "dense-pile-fused", "multi-collider", "batched-worlds", "boundary-shell-auto", "full-size-cfg2")]


@pytest.mark.parametrize("tune", ["12", "13"], ids=["tune12-128-record-tiles", "tune13-tma-bulk-windows"])
@pytest.mark.parametrize("knobs", [{}, FORCED], ids=["default", "forced-pool-crowded"])
@pytest.mark.parametrize("case", SMALL_TILE_CASES, ids=[c[0] for c in SMALL_TILE_CASES])
def test_tile_kernel_variants_on_gpu(case, knobs, tune, monkeypatch):
    """BLOBS_PARAM_TUNE 12: the same kernel with 128-record tiles (128-thread CTAs); 13: windows fetched by TMA bulk copies
    (cp.async.bulk + mbarrier; this one cannot run on the host-compiled build at all - the GPU run is its first). The very last
    tests of the suite for that reason."""
    import os

    if tune == "13" and os.environ.get("BLOBS_TEST_EXPERIMENTAL") != "1":
        pytest.skip("inline-PTX path that has never executed anywhere: run it on purpose (BLOBS_TEST_EXPERIMENTAL=1, under a timeout)")
    name, fn, kw = case
    monkeypatch.setenv("BLOBS_B200_TUNE", tune)
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    fn(**kw)
