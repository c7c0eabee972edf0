"""Host-side checks of bench.py that need no GPU: the argument table, the traffic look-up rule, the reference arm's JSON shape.
(The measured paths themselves run on the GPU box; the strip frame loop and the N > 1 self-check of bench.py are exercised by the
emulated-rank cases of test_multi_gpu.py.)"""
import json
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench_source():
    return open(os.path.join(REPO, "bench.py")).read()


def test_every_argument_bench_reads_is_declared():
    src = _bench_source()
    used = set(re.findall(r"\bargs\.([a-z_]+)", src))
    declared = {a.replace("-", "_") for a in re.findall(r'add_argument\("--([a-z-]+)"', src)}
    assert used <= declared, f"bench.py reads undeclared arguments: {sorted(used - declared)}"


def test_traffic_is_only_reported_for_the_same_kernel_workload_and_size():
    sys.path.insert(0, REPO)
    import bench

    cap = json.load(open(os.path.join(REPO, "profiles", "kernel_traffic.json")))["captures"]
    assert cap, "profiles/kernel_traffic.json holds no capture"
    for e in cap:
        assert bench.ncu_traffic(e["kernel"], e["colliders"], e["workload"]) == float(e["dram_bytes_per_launch"])
        assert bench.ncu_traffic(e["kernel"], e["colliders"] + 1, e["workload"]) is None
        assert bench.ncu_traffic(e["kernel"], e["colliders"], "cfg3") is None
        assert bench.ncu_traffic("k_other", e["colliders"], e["workload"]) is None


def test_reference_arm_prints_one_json_line_on_cpu():
    """`bench.py --impl reference` times the CPU restatement of the reference algorithm (no GPU needed): one JSON line with the
    contract's keys, `impl` = reference, zero host<->device bytes."""
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "sphere_steps_per_sec" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
