"""Multi-GPU path. GPU part: strip-decomposed world == single-GPU world, bit for bit (needs >= 2 GPUs, run
with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`). CPU part: the host-side partition logic under
world_size-2 gloo."""
import os
import subprocess
import sys

import numpy as np
import pytest


def _free_port():
    """a TCP port nobody listens on right now (fixed ports collide when two test runs share a machine)"""
    import socket

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_strip_world_matches_single_gpu():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(REPO, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0


@pytest.mark.gpu
def test_strip_world_peer_memory_exchange_matches_single_gpu():
    """Same, with the per-substep exchange done by k_strip_push (peer-memory stores through CUDA IPC mappings, BLOBS_PARAM_STRIP_P2P)
    instead of ncclSend/ncclRecv. The worker asserts that the peer path is really active on every rank."""
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else 4
    env = dict(os.environ, BLOBS_B200_STRIP_P2P="1", STRIP_TEST_EXPECT_P2P="1", STRIP_TEST_IO="pipelined", BLOBS_B200_LIST="0", STRIP_TEST_EXPECT_LISTS="0")   # + bench.py's pipelined frame loop
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(REPO, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0


@pytest.mark.gpu
def test_strip_world_neighbour_lists_match_single_gpu():
    """The neighbour-list pipeline on strips (the library default when the peer mappings are available): k_step stores the ghost
    records straight into the neighbours' arrays, every rank takes the same rebuild decision from all ranks' numbers, bodies
    change owner at rebuilds only. Merged strips == the single-GPU world, bit for bit."""
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else 4
    env = dict(os.environ, BLOBS_B200_STRIP_P2P="1", STRIP_TEST_EXPECT_P2P="1", BLOBS_B200_LIST="1", STRIP_TEST_EXPECT_LISTS="1",
               STRIP_TEST_BENCH_PROBE="1")   # + bench.py's own N > 1 self-check (`parity_vs_single_gpu`)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(REPO, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0


_FORCED = {"BLOBS_B200_POOL": "1", "BLOBS_B200_POOL_MIN": "1", "BLOBS_B200_CROWDED": "1"}
_SHELL = {"STRIP_TEST_SCENE": "shell", "STRIP_TEST_STEPS": "8"}
_P2P = {"BLOBS_B200_STRIP_P2P": "1", "STRIP_TEST_EXPECT_P2P": "1"}
# the combinations that add nothing new to the ones below only run with BLOBS_TEST_SLOW=1 (each costs ~40 s of emulation)
_slow = pytest.mark.skipif(os.environ.get("BLOBS_TEST_SLOW") != "1", reason="redundant combination; set BLOBS_TEST_SLOW=1")
_LISTS = {"BLOBS_B200_LIST": "1", "STRIP_TEST_EXPECT_LISTS": "1"}
_STRIP_CASES = [
    pytest.param({}, id="gas-default", marks=_slow),   # (gas-pipelined-host-io below is the same configuration + the pipelined frame loop)
    pytest.param(dict(_FORCED), id="gas-forced-pool-crowded", marks=_slow),
    pytest.param({**_SHELL, **_FORCED}, id="shell-forced-pool-crowded"),
    pytest.param(dict(_SHELL), id="shell-default", marks=_slow),
    # neighbour lists on strips: ghost records stored straight into the neighbour's arrays by k_step, rebuild decision combined
    # over all ranks, migration at rebuilds only (needs the peer-memory mappings)
    pytest.param({**_P2P, **_LISTS, "STRIP_TEST_BENCH_PROBE": "1"}, id="gas-p2p-lists+bench-probe"),   # + bench.py's N > 1 self-check: a second strip world in the same process
    pytest.param({**_P2P, **_LISTS, "STRIP_TEST_RANKS": "1", "STRIP_TEST_STEPS": "10"}, id="one-rank-strip-lists"),   # the strip code paths with no peer (profiles/r2_scripts/strip_diag.py)
    pytest.param({**_SHELL, **_P2P, **_LISTS, "BLOBS_B200_CROWDED": "1", "STRIP_TEST_RANKS": "3"}, id="shell-p2p-3ranks-lists-crowded"),
    pytest.param({**_P2P, **_LISTS, "STRIP_TEST_IO": "pipelined", "STRIP_TEST_STEPS": "12"}, id="gas-p2p-lists-pipelined-host-io", marks=_slow),
    # automatic mode on an agitated scene: the ranks start on lists, see them rebuilt every substep and fall back to the grid
    # pipeline together, in the middle of the run
    pytest.param({**_SHELL, **_P2P, "BLOBS_B200_LIST": "2", "STRIP_TEST_EXPECT_LISTS": "1", "STRIP_TEST_EXPECT_LISTS_AFTER": "0", "STRIP_TEST_STEPS": "10"}, id="shell-p2p-auto-falls-back-to-grid"),
    # the cell-grid pipeline with the peer-memory exchange (k_strip_push every substep)
    pytest.param({**_P2P, "BLOBS_B200_LIST": "0", "STRIP_TEST_EXPECT_LISTS": "0"}, id="gas-p2p-grid", marks=_slow),
    pytest.param({**_SHELL, **_P2P, **_FORCED, "BLOBS_B200_LIST": "0", "STRIP_TEST_RANKS": "3"}, id="shell-p2p-3ranks-grid-forced", marks=_slow),
    pytest.param({"STRIP_TEST_IO": "pipelined", "STRIP_TEST_STEPS": "12"}, id="gas-pipelined-host-io"),
]


@pytest.mark.emu
@pytest.mark.parametrize("knobs", _STRIP_CASES)
def test_strip_world_matches_single_world_emulated_ranks(knobs):
    """The same worker without GPUs: 2 CPU processes, each running the host-compiled build of the CUDA sources (tests/emu),
    exchanging ghosts and migrants every substep through a socket stand-in for NCCL; merged result == single world, bit for
    bit, with bodies migrating between the strips."""
    from .emu_loader import build

    build()
    env = dict(os.environ, BLOBS_TEST_EMU="1", STRIP_TEST_SIDE="48", STRIP_TEST_STEPS="24")
    env.update(knobs)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={env.get('STRIP_TEST_RANKS', '2')}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(REPO, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert "mismatches=0" in r.stdout


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, REPO)
    from blobs_b200 import scenes as S
    from blobs_b200 import strips

    sc = S.cfg2(seed=3, side=64)
    x = sc.bodies["position"]["x"]
    half = len(x) // 2
    local = x[:half] if rank == 0 else x[half:]      # each rank only looks at part of the scene
    edges = strips.agree_edges(dist, local, world)
    owner = strips.owner_of(x, edges)
    q.put((rank, edges.tobytes(), owner.tobytes()))
    dist.destroy_process_group()


def test_strip_partition_agrees_across_ranks_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert got[0][1] == got[1][1], "ranks disagree on the strip edges"
    assert got[0][2] == got[1][2], "ranks disagree on ownership"
    owner = np.frombuffer(got[0][2], dtype=np.int32)
    assert set(np.unique(owner)) == {0, 1} and abs(int((owner == 0).sum()) - int((owner == 1).sum())) < len(owner) * 0.1


def test_strip_edges_properties():
    sys.path.insert(0, REPO)
    from blobs_b200 import strips

    e = strips.strip_edges(-537.6, 537.6, 8)
    assert e.dtype == np.float32 and len(e) == 9 and np.isneginf(e[0]) and np.isposinf(e[-1])
    x = np.array([-1e9, -537.6, e[1], np.nextafter(e[1], np.float32(-np.inf)), 0.0, 537.6, 1e9, np.nan], dtype=np.float32)
    o = strips.owner_of(x, e)
    assert o.tolist() == [0, 0, 1, 0, 4, 7, 7, 0]
    assert strips.check_strip_width(e, 0.5) and not strips.check_strip_width(strips.strip_edges(0, 10, 8), 0.5)
