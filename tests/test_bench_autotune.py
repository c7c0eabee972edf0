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
