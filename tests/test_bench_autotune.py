"""bench.py's kernel-variant autotuner: the probe (k_main vs k_tile on the same scene, bit-exact parity required) run against the
host-compiled build of the kernels, and the decision logic around the child process."""
import argparse
import json
import os
import subprocess
import sys
import types

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import bench  # noqa: E402

from .emu_loader import emulated  # noqa: E402


def _args(**kw):
    d = dict(workload="cfg1", warmup=3, steps=6, device=0, tune=0, probe=True)
    d.update(kw)
    return argparse.Namespace(**d)


@pytest.mark.emu
def test_probe_compares_the_two_variants_bit_for_bit(monkeypatch, capsys):
    import torch

    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    with emulated():
        bench.run_probe(_args())
    line = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")][-1]
    p = json.loads(line)
    assert p["probe"] is True and p["parity"] == {"11": True, "12": True} and p["steps"] == 6
    assert all(p["ms"][k] > 0 for k in ("0", "11", "12"))


@pytest.mark.parametrize("stdout,rc,want_tune,want_key", [
    ('{"probe": true, "ms": {"0": 2.0, "11": 1.5, "12": 1.7}, "parity": {"11": true, "12": true}, "steps": 30, "warmup": 200}\n', 0, 11, "k_tile"),
    ('{"probe": true, "ms": {"0": 2.0, "11": 1.5, "12": 1.2}, "parity": {"11": true, "12": true}, "steps": 30, "warmup": 200}\n', 0, 12, "k_tile<128>"),
    ('{"probe": true, "ms": {"0": 2.0, "11": 1.99, "12": 2.4}, "parity": {"11": true, "12": true}, "steps": 30, "warmup": 200}\n', 0, 0, "k_main"),    # not 3 % faster
    ('{"probe": true, "ms": {"0": 2.0, "11": 1.0, "12": 1.8}, "parity": {"11": false, "12": true}, "steps": 30, "warmup": 200}\n', 0, 12, "k_tile<128>"),  # faster but wrong: never
    ('{"probe": true, "ms": {"0": 2.0, "11": 1.0, "12": 1.0}, "parity": {"11": false, "12": false}, "steps": 30, "warmup": 200}\n', 0, 0, "k_main"),
    ("", 1, 0, None),                                                                                                      # the child died
    ("garbage\n", 0, 0, None),
])
def test_autotune_decision(monkeypatch, stdout, rc, want_tune, want_key):
    monkeypatch.delenv("BLOBS_BENCH_AUTOTUNE", raising=False)
    monkeypatch.setattr(subprocess, "run", lambda *a, **k: types.SimpleNamespace(returncode=rc, stdout=stdout, stderr="boom"))
    tune, rep = bench.autotune_main_kernel(_args(workload="cfg2", warmup=200, steps=100, probe=False), 0)
    assert tune == want_tune and rep.get("chosen") == want_key
    if want_key is None:
        assert "probe failed" in rep["result"]


def test_autotune_off_when_a_variant_is_forced_or_the_child_hangs(monkeypatch):
    tune, rep = bench.autotune_main_kernel(_args(tune=9, probe=False), 0)
    assert tune == 9 and rep["mode"].startswith("off")

    def hang(*a, **k):
        raise subprocess.TimeoutExpired(cmd="bench.py --probe", timeout=1)

    monkeypatch.setattr(subprocess, "run", hang)
    tune, rep = bench.autotune_main_kernel(_args(probe=False), 0)
    assert tune == 0 and rep["result"] == "probe timed out"


@pytest.mark.emu
def test_strip_probe_children_form_their_own_group():
    """N > 1: two parent ranks under torchrun each start one child; the children (host-compiled kernels, gloo, socket stand-in for
    NCCL) rendezvous on MASTER_PORT + 23, run the small strip world through k_main and k_tile and agree on one verdict."""
    import socket

    from .emu_loader import build

    build()
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    env = dict(os.environ, BLOBS_TEST_EMU="1", BLOBS_BENCH_PROBE_TIMEOUT="400")
    env.pop("BLOBS_BENCH_AUTOTUNE", None)
    env.pop("BLOBS_B200_STRIP_P2P", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(REPO, "tests", "strip_autotune_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    # the ranks share one pipe: tolerate two verdicts on one line
    verdicts = [json.loads(c) for l in r.stdout.splitlines() for c in l.split("VERDICT ")[1:]]
    assert sorted(v["rank"] for v in verdicts) == [0, 1]
    for v in verdicts:
        rep = v["report"]
        assert rep.get("parity_bit_exact") == {"k_tile + ncclSend/ncclRecv": True, "k_main + peer-memory exchange": True, "k_tile + peer-memory exchange": True}, rep
        assert rep["chosen"] in bench.STRIP_VARIANT_NAME.values() and v["tune"] in (0, 11) and v["p2p"] in (0, 1)
        assert (v["p2p"] == 1) == rep["chosen"].endswith("peer-memory exchange") and (v["tune"] == 11) == rep["chosen"].startswith("k_tile")
    assert verdicts[0]["report"]["ms_per_step"] == verdicts[1]["report"]["ms_per_step"]   # all-reduced: every rank decides alike
