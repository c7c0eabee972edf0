#!/usr/bin/env python
"""bench.py — sphere-steps/sec of the blobs::Physics::step hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg2]

A "step" is one Physics::step(1/60) (8 substeps) over a synthetic scene. At N=1 the workload is BASELINE
config #2 (1 048 576 single-collider spheres in one world). Prints ONE JSON line (rank 0).

Timing: `value` is device time (CUDA events recorded by the library on its own stream around each step's
kernels), inputs resident in HBM, max over ranks; between timed steps a 256 MiB buffer is rewritten to flush L2
(outside the per-step events). `e2e` is wall-clock through the public C ABI with per-step pinned-host -> device forces
and device -> pinned-host positions inside the timed region (N = 1: the ABI's pipelined host I/O, copies on their own
streams; `e2e.sync_value` = the blocking calls on a short sample).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line (NCCL prints its version banner to stdout otherwise)

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "sphere_steps_per_sec"
UNIT = "sphere-steps/s"
DT = 1.0 / 60.0
# Algorithmic bytes per collider-substep
B_MAIN = 64 + 52      # fused contact + verlet + snapshot + clamp + key kernel (dominant)
B_PIPELINE = 236      # whole substep


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg2_varied", "cfg2_dense", "cfg3", "cfg4", "cfg2_4m", "cfg2_16m"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--tune", type=int, default=0, help="kernel variant selector (BLOBS_PARAM_TUNE)")
    ap.add_argument("--list", type=int, default=None, help="BLOBS_PARAM_LIST: 0 = cell grid every substep, 1 = neighbour lists")
    ap.add_argument("--skin", type=float, default=None, help="BLOBS_PARAM_SKIN (fraction of the largest radius)")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--probe", action="store_true", help="(internal) kernel-variant probe run by the autotuner in a child process")
    ap.add_argument("--device", type=int, default=None, help="(internal) CUDA device of the probe")
    ap.add_argument("--probe-strips", action="store_true", help="(internal) strip-mode kernel-variant probe: one child per rank, own process group")
    return ap.parse_args()


def make_scene(name, seed=1):
    from blobs_b200 import scenes as S

    if name == "cfg1":
        return S.cfg1(seed), "cfg1: 1024 spheres r~U[0.05,0.2) in a circle constraint R=8"
    if name == "cfg2":
        return S.cfg2(seed), "cfg2: 1048576 single-collider spheres r=0.5, jittered lattice pitch 1.05, circle constraint R=800, g=(0,-30)"
    if name == "cfg2_varied":
        return S.cfg2(seed, varied=True), "cfg2 variant: 1048576 spheres r~U[0.25,0.5)"
    if name == "cfg2_dense":
        return S.cfg2_dense(seed, side=1024), "cfg2 dense variant: 1048576 spheres r=0.5 on a pitch-0.9 lattice (every sphere starts overlapping its 4 neighbours)"
    if name == "cfg2_4m":
        return S.cfg2(seed, side=2048), "cfg2 scaled: 4194304 spheres"
    if name == "cfg2_16m":
        return S.cfg2(seed, side=4096), "cfg5 single-GPU form: 16777216 spheres in one world"
    if name == "cfg4":
        return S.cfg4(100_000, 16, seed), "cfg4: 100000 soft blobs x 16 bodies, fixed joints + springs"
    raise ValueError(name)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.idx)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def mark(self):
        """the timed region starts here: remember how many samples belong to the warm-up tail"""
        self.f.flush()
        try:
            self.n_before = sum(1 for _ in open(self.f.name))
        except OSError:
            self.n_before = 0

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for line in self.f.read().strip().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "samples_before_timed_region": getattr(self, "n_before", 0), "interval_ms": 20, "power_w_max": max(power)}


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu summary, if one exists."""
    p = os.path.join(REPO, "profiles", "k_main_traffic.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_launch"])
        except Exception:
            return None
    return None


def cpu_reference_sample(budget_s, steps_hint=None, warmup=0):
    """Times the reference's own algorithm (brute-force O(C^2) Physics::step, CPU oracle = faithful restatement, 1 thread —
    the reference is !Send) on a bounded sample of cfg2: a side x side sub-block of the same lattice."""
    from blobs_b200 import scenes as S
    from oracle import oracle_py

    def run(side, steps, wu):
        sc = S.cfg2(seed=1, side=side)
        o = oracle_py.OracleWorld(gravity=sc.gravity, maintain_spatial_hash=True, record_events=True)
        S.build(o, sc)
        if wu:
            o.step_n_timed(DT, wu)
            o.events_drain(); o.pairs_drain()
        secs = o.step_n_timed(DT, steps)
        return sc.n_colliders, secs

    # calibrate the pair-test rate on a tiny block
    n0, s0 = run(32, 2, 0)
    rate = (n0 * n0 / 2.0) * 8 * 2 / max(s0, 1e-6)  # pair tests / s
    if steps_hint is None:
        side, steps = 96, 1
        per_step = (side * side) ** 2 / 2.0 * 8 / rate
        steps = max(1, int(budget_s / max(per_step, 1e-6)))
        steps = min(steps, 50)
    else:
        steps = steps_hint
        per_step_budget = budget_s / max(1, steps + warmup)
        c = (2.0 * per_step_budget * rate / 8.0) ** 0.5
        side = int(max(16, min(128, c ** 0.5)))
    n, secs = run(side, steps, warmup)
    full = 1048576
    return {
        "value": n * steps / secs, "unit": UNIT, "cores": 1, "kind": "port",
        "sample": f"{n} spheres ({side}x{side} sub-block of the cfg2 lattice) x {steps} Physics::step(1/60), brute-force O(C^2) pair loop as in the "
                  f"reference (physics.rs:241-317), single thread, {secs:.2f} s; host has {os.cpu_count()} cores; the reference cannot be built here "
                  f"(Rust), so this is the C++ oracle restatement. O(C^2): at the full {full} spheres the same loop would run ~{full / n:.0f}x slower per sphere",
        "pair_tests_per_s": rate,
    }, n, steps, secs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, n, steps, secs = cpu_reference_sample(120.0, steps_hint=args.steps, warmup=args.warmup)
    v = cb["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2 (bounded sample): " + cb["sample"]},
        "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def build_single_world(workload, device):
    """one world of `workload` on `device` (the N = 1 form of every workload)"""
    import blobs_b200
    from blobs_b200 import scenes as S

    if workload == "cfg3":
        worlds = [S.cfg1(1 + wid, n_side=16) for wid in range(4096)]
        w = blobs_b200.World(gravity=worlds[0].gravity, device=device, body_capacity=256 * 4096, collider_capacity=256 * 4096)
        S.build_batch(w, worlds)
        return w, 256 * 4096
    sc, _ = make_scene(workload, seed=1)
    w = blobs_b200.World(gravity=sc.gravity, device=device, body_capacity=sc.n_bodies, collider_capacity=sc.n_colliders)
    S.build(w, sc)
    return w, sc.n_bodies


_EMU_CTX = None
STRIP_VARIANTS = ((0, 0), (11, 0), (0, 1), (11, 1))   # (BLOBS_PARAM_TUNE, BLOBS_PARAM_STRIP_P2P) probed at N > 1; the first is the baseline
STRIP_VARIANT_NAME = {"0": "k_main + ncclSend/ncclRecv", "11": "k_tile + ncclSend/ncclRecv", "0+p2p": "k_main + peer-memory exchange", "11+p2p": "k_tile + peer-memory exchange"}
PROBE_VARIANTS = (0, 11, 12)   # BLOBS_PARAM_TUNE: k_main, k_tile (256-record tiles), k_tile (128-record tiles)
VARIANT_NAME = {0: "k_main", 11: "k_tile", 12: "k_tile<128>"}


def run_probe(args):
    """Child process of the autotuner: the same scene once per variant (k_main = BLOBS_PARAM_TUNE 0, k_tile = 11, k_tile with 128-record
    tiles = 12), W warm-up steps each, then K steps timed in alternating blocks of 5. Prints {"ms": {tune: ..}, "parity": {tune: bool}}:
    parity = positions, previous positions and velocities are bit-identical to k_main's after all W + K steps (they must be: the
    variants only differ in how threads are mapped onto the same arithmetic). Runs in its own process so that a fault in the
    not-yet-measured variant cannot take the benchmark down with it."""
    import numpy as np
    import torch

    import blobs_b200

    dev = args.device if args.device is not None else 0
    torch.cuda.set_device(dev)
    worlds = {}
    for tune in PROBE_VARIANTS:
        w, _ = build_single_world(args.workload, dev)
        w.set_param(blobs_b200.abi.PARAM_TUNE, tune)
        w.step(DT, n=max(args.warmup, 1))
        worlds[tune] = w
    ms = {t: 0.0 for t in PROBE_VARIANTS}
    done = 0
    while done < args.steps:
        blk = min(5, args.steps - done)
        for tune in PROBE_VARIANTS:
            for _ in range(blk):
                ms[tune] += worlds[tune].step(DT)["gpu_ms"]
        done += blk
    a, _ = worlds[0].download_bodies()
    parity = {}
    for tune in PROBE_VARIANTS[1:]:
        b, _ = worlds[tune].download_bodies()
        parity[str(tune)] = bool(all(np.array_equal(a[f][c].view(np.uint32), b[f][c].view(np.uint32))
                                     for f in ("position", "position_old", "calculated_velocity") for c in ("x", "y")))
    print(json.dumps({"probe": True, "ms": {str(k): v / max(args.steps, 1) for k, v in ms.items()}, "parity": parity, "steps": args.steps, "warmup": args.warmup}), flush=True)


def autotune_main_kernel(args, device):
    """Picks the dominant kernel's variant for this run by MEASUREMENT, like a library autotuner: a child process (run_probe) times
    k_main against k_tile on this GPU, in the regime the timed window sits in, and checks that they agree bit for bit. k_tile is
    used only if the child finished cleanly, parity held and it was at least 3 % faster. Returns (tune, report)."""
    if args.tune or os.environ.get("BLOBS_BENCH_AUTOTUNE", "1") == "0":
        return args.tune, {"mode": "off (variant forced)" if args.tune else "off"}
    cmd = [sys.executable, os.path.abspath(__file__), "--probe", "--workload", args.workload, "--warmup", str(max(args.warmup, 3)), "--steps", str(min(max(args.steps, 10), 30)),
           "--device", str(device)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    rep = {"mode": "probe in a child process: k_main (tune 0) vs k_tile (tune 11 / 12), same scene and window, bit-exact parity required"}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=float(os.environ.get("BLOBS_BENCH_PROBE_TIMEOUT", "180")), env=env)
        line = next((l for l in r.stdout.splitlines() if l.startswith("{") and '"probe"' in l), None)
        if r.returncode != 0 or line is None:
            rep["result"] = f"probe failed (rc={r.returncode}): {r.stderr.strip().splitlines()[-1] if r.stderr.strip() else 'no output'}"[:300]
            return 0, rep
        p = json.loads(line)
        rep.update({"ms_per_step": {VARIANT_NAME[int(k)]: v for k, v in p["ms"].items()},
                    "parity_bit_exact": {VARIANT_NAME[int(k)]: v for k, v in p["parity"].items()}, "probe_steps": p["steps"]})
        best, best_ms = 0, 0.97 * p["ms"]["0"]
        for k, v in p["ms"].items():
            if int(k) and p["parity"].get(k) is True and v < best_ms:
                best, best_ms = int(k), v
        rep["chosen"] = VARIANT_NAME[best]
        return best, rep
    except subprocess.TimeoutExpired:
        rep["result"] = "probe timed out"
        return 0, rep
    except Exception as e:  # noqa: BLE001 - the probe is optional
        rep["result"] = f"probe error: {e}"[:300]
        return 0, rep


def run_probe_strips(args):
    """Child processes of the N > 1 autotuner (one per rank, their own process group on another port): a small strip-decomposed
    world (128 lattice columns per rank) once through k_main and once through k_tile, W warm-up + K timed steps each. Prints
    {"ms": {tune: max over ranks}, "parity": {tune: bool over all ranks}}: parity = every rank ends up owning the same bodies with
    bit-identical positions in both runs. BLOBS_TEST_EMU=1 runs the same thing on the host-compiled build over gloo (tests)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    emu = os.environ.get("BLOBS_TEST_EMU") == "1"
    if emu:   # test infrastructure: CPU rank processes on the host-compiled kernels, socket stand-in for NCCL (tests/emu)
        sys.path.insert(0, os.path.join(REPO, "tests"))
        import emu_loader

        os.environ["BLOBS_EMU_NCCL_LIB"] = os.path.join(emu_loader.EMU_DIR, "libnccl_fake.so")
        global _EMU_CTX
        _EMU_CTX = emu_loader.emulated()   # kept alive for the life of the process
        _EMU_CTX.__enter__()
        dist.init_process_group("gloo")
        dev = "cpu"
        nx, ny = 24 * world, 48
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dev = "cuda"
        nx, ny = 128 * world, 2048
    import blobs_b200
    from blobs_b200 import scenes as S
    from blobs_b200 import strips

    sc = S.lattice_scene(nx, ny, 1.05, (0.0, 0.0), 1, 0.5, 0.5, jitter=0.04, vel_disc=1.0, constraint_r=0.8 * max(nx, ny), name="cfg5-probe", cell_size=1.0)
    edges = strips.strip_edges(float(sc.bodies["position"]["x"].min()), float(sc.bodies["position"]["x"].max()), world)
    # (BLOBS_PARAM_TUNE, peer-memory exchange): the baseline first. A peer exchange that delivered stale ghosts, or waited out its
    # timeout (bit 3 of nan_detected), fails the parity check like any other wrong variant.
    variants = STRIP_VARIANTS
    ms, owned, pos, flags, active = {}, {}, {}, {}, {}
    for v in variants:
        tune, p2p = v
        w = blobs_b200.World(gravity=sc.gravity, device=local, body_capacity=sc.n_bodies, collider_capacity=sc.n_colliders)
        S.build(w, sc)
        w.set_param(blobs_b200.abi.PARAM_TUNE, tune)
        w.set_param(blobs_b200.abi.PARAM_STRIP_P2P, p2p)
        uid = torch.from_numpy(blobs_b200.World.strip_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).to(dev)
        dist.broadcast(uid, 0)
        w.strip_configure(rank, world, float(edges[rank]), float(edges[rank + 1]), uid.cpu().numpy(), ghost_capacity=4 * ny, migrate_capacity=2 * ny)
        active[v] = int(w.get_param(blobs_b200.abi.PARAM_STRIP_P2P)) == p2p   # a requested peer exchange may have fallen back to NCCL
        flags[v] = w.step(DT, n=max(args.warmup, 1))["nan_detected"]
        t = 0.0
        for _ in range(args.steps):
            st = w.step(DT)
            t += st["gpu_ms"]
            flags[v] |= st["nan_detected"]
        ms[v] = t / max(args.steps, 1)
        owned[v] = w.strip_owned().astype(bool)
        pos[v] = w.read_positions()
        del w
    base = variants[0]
    same = {}
    for v in variants[1:]:
        same[v] = bool(active[v] and flags[v] == 0 and flags[base] == 0 and np.array_equal(owned[base], owned[v])
                       and np.array_equal(pos[base][owned[base]].view(np.uint32), pos[v][owned[v]].view(np.uint32)))
    t_ms = torch.tensor([ms[v] for v in variants], dtype=torch.float64, device=dev)
    t_ok = torch.tensor([1.0 if same[v] else 0.0 for v in variants[1:]], dtype=torch.float64, device=dev)
    dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
    key = lambda v: f"{v[0]}+p2p" if v[1] else str(v[0])
    print(json.dumps({"probe": True, "ms": {key(v): float(t_ms[i]) for i, v in enumerate(variants)},
                      "parity": {key(v): bool(float(t_ok[i]) > 0.5) for i, v in enumerate(variants[1:])}, "steps": args.steps, "warmup": args.warmup,
                      "spheres_per_rank": nx * ny // world}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def autotune_strips(args):
    """N > 1 (strip-decomposed world): every rank starts ONE child (run_probe_strips); the children form their own process group
    on MASTER_PORT + 23 and time k_main against k_tile on a small strip world, with bit-exact parity required on every rank. All
    ranks read the same all-reduced verdict. Any failure (a child dies, the group hangs until the timeout) means k_main."""
    forced_p2p = 1 if os.environ.get("BLOBS_B200_STRIP_P2P", "0") not in ("", "0") else 0
    if args.tune or forced_p2p or os.environ.get("BLOBS_BENCH_AUTOTUNE", "1") == "0":
        return args.tune, forced_p2p, {"mode": "off (variant forced)" if (args.tune or forced_p2p) else "off"}
    env = {k: v for k, v in os.environ.items() if not k.startswith("TORCHELASTIC_")}   # the children rendezvous among themselves (rank 0 hosts the store)
    env["MASTER_PORT"] = str(int(os.environ.get("MASTER_PORT", "29500")) + 23)
    env.setdefault("MASTER_ADDR", "127.0.0.1")
    cmd = [sys.executable, os.path.abspath(__file__), "--probe-strips", "--warmup", "30", "--steps", "20"]
    rep = {"mode": "strip probe, one child process per rank in their own process group: {k_main, k_tile} x {ncclSend/ncclRecv, peer-memory exchange}, "
                   "bit-exact parity with the baseline required on every rank"}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=float(os.environ.get("BLOBS_BENCH_PROBE_TIMEOUT", "180")), env=env)
        line = next((l for l in r.stdout.splitlines() if l.startswith("{") and '"probe"' in l), None)
        if r.returncode != 0 or line is None:
            rep["result"] = f"probe failed (rc={r.returncode}): {r.stderr.strip().splitlines()[-1] if r.stderr.strip() else 'no output'}"[:300]
            return 0, 0, rep
        p = json.loads(line)
        rep.update({"ms_per_step": {STRIP_VARIANT_NAME.get(k, k): v for k, v in p["ms"].items()},
                    "parity_bit_exact": {STRIP_VARIANT_NAME.get(k, k): v for k, v in p["parity"].items()}, "probe_steps": p["steps"],
                    "probe_spheres_per_rank": p.get("spheres_per_rank")})
        best, best_ms = "0", 0.97 * p["ms"]["0"]
        for k, v in p["ms"].items():
            if k != "0" and p["parity"].get(k) is True and v < best_ms:
                best, best_ms = k, v
        rep["chosen"] = STRIP_VARIANT_NAME.get(best, best)
        return int(best.split("+")[0]), (1 if best.endswith("+p2p") else 0), rep
    except subprocess.TimeoutExpired:
        rep["result"] = "probe timed out"
        return 0, 0, rep
    except Exception as e:  # noqa: BLE001 - the probe is optional
        rep["result"] = f"probe error: {e}"[:300]
        return 0, 0, rep


def strip_pipelined_loop(w, K, forces, sl, xy, cnt, io_cap, on_step=None):
    """K frames of a strip-decomposed world with the pipelined distributed host I/O of the C ABI. `sl`, `xy`, `cnt` are pairs of
    pinned host tensors (slot list, positions, count); both slot lists / counts hold the current owned list on entry. The newest
    list that has landed is in pair (K - 1) & 1 on return. Returns the bytes moved. (tests/multi_gpu_worker.py runs this very loop.)"""
    io_bytes = 0
    newest = 1                                                    # pair holding the newest list that has landed (frame 0 writes pair 0)
    m = min(int(cnt[newest][0]), io_cap)
    w.forces_indexed_upload_async_ptr(sl[newest].data_ptr(), forces.data_ptr(), m)
    io_bytes += m * 12
    for i in range(K):
        w.apply_forces_indexed_uploaded()
        if i + 1 < K:
            m = min(int(cnt[newest][0]), io_cap)
            w.forces_indexed_upload_async_ptr(sl[newest].data_ptr(), forces.data_ptr(), m)   # frame i+1's forces travel under frame i's kernels
            io_bytes += m * 12
        st = w.step(DT)
        if on_step is not None:
            on_step(st)
        w.io_sync()                                               # frame i-1's list has landed in pair (i-1)&1; the upload has left pair `newest`
        if i:
            newest = (i - 1) & 1
        w.read_owned_positions_async_ptr(sl[i & 1].data_ptr(), xy[i & 1].data_ptr(), cnt[i & 1].data_ptr(), io_cap)   # under frame i+1's kernels
        io_bytes += min(int(cnt[newest][0]), io_cap) * 12         # (the copy moves the owned-list bound; counted as the valid part)
    w.io_sync()
    return io_bytes


def run_ours(args):
    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; blobs_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import blobs_b200
    from blobs_b200 import scenes as S

    # which variant of the dominant kernel runs: measured first, in child processes (a fault in a variant that had never run on a
    # GPU when it was committed costs the probe, not the benchmark)
    tune_report = {"mode": "off (N > 1, independent worlds)"}
    if world == 1:
        args.tune, tune_report = autotune_main_kernel(args, local)
    strip_p2p = 0
    if world > 1 and args.workload == "cfg2":   # strip-decomposed world: probed by a group of child processes, one per rank
        args.tune, strip_p2p, tune_report = autotune_strips(args)

    scaling = "weak"
    if args.workload == "cfg3":
        # BASELINE config #3: 4096 independent worlds x 256 bodies, block-partitioned over the ranks, no collective (strong scaling)
        n_worlds_total = 4096
        lo, hi = rank * n_worlds_total // world, (rank + 1) * n_worlds_total // world
        worlds = [S.cfg1(1 + wid, n_side=16) for wid in range(lo, hi)]
        w = blobs_b200.World(gravity=worlds[0].gravity, device=local, body_capacity=256 * (hi - lo), collider_capacity=256 * (hi - lo))
        S.build_batch(w, worlds)
        n = nb = 256 * (hi - lo)
        desc = f"cfg3: {n_worlds_total} batched independent worlds x 256 bodies (cfg1 at 16x16, circle R=4), worlds {lo}..{hi - 1} on this rank"
        scaling = "strong"
    elif world > 1 and args.workload == "cfg2":
        # BASELINE config #5 family: ONE world of 512*N x 4096 spheres (N=8: 16 777 216), cut into N vertical strips of 512 lattice
        # columns (2 097 152 spheres per GPU, weak scaling); per-substep ghost + migration exchange with both neighbours over NCCL.
        from blobs_b200 import strips

        nx, ny = 512 * world, 4096
        sc = S.lattice_scene(nx, ny, 1.05, (0.0, 0.0), 1, 0.5, 0.5, jitter=0.04, vel_disc=1.0, constraint_r=0.8 * max(nx, ny), name="cfg5", cell_size=1.0)
        desc = (f"cfg5 family: one world of {nx}x{ny} = {nx * ny} spheres r=0.5 (pitch 1.05, circle R={0.8 * max(nx, ny):.0f}), strip-decomposed over {world} GPUs "
                f"({nx * ny // world} spheres per GPU), ghost/migration exchange with both neighbours every substep")
        w = blobs_b200.World(gravity=sc.gravity, device=local, body_capacity=sc.n_bodies, collider_capacity=sc.n_colliders)
        S.build(w, sc)
        if strip_p2p:
            w.set_param(blobs_b200.abi.PARAM_STRIP_P2P, 1)   # chosen by the probe (BLOBS_B200_STRIP_P2P=1 in the environment forces it)
        edges = strips.strip_edges(float(sc.bodies["position"]["x"].min()), float(sc.bodies["position"]["x"].max()), world)
        uid = torch.from_numpy(blobs_b200.World.strip_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).cuda()
        dist.broadcast(uid, 0)
        w.strip_configure(rank, world, float(edges[rank]), float(edges[rank + 1]), uid.cpu().numpy(), ghost_capacity=4 * ny, migrate_capacity=2 * ny)   # bounds: the band next to an edge holds ~1-2 lattice columns (ny rows each)
        n = int(w.strip_owned().sum())
        nb = sc.n_bodies
        del sc
    else:
        # other workloads at N > 1: one independent world per rank (no data-path collective; "replicas", weak scaling).
        sc, desc = make_scene(args.workload, seed=1 + rank)
        if world > 1:
            desc += f"; one independent world per GPU ({world} replicas, no collective)"
        w = blobs_b200.World(gravity=sc.gravity, device=local, body_capacity=sc.n_bodies, collider_capacity=sc.n_colliders)
        S.build(w, sc)
        n = sc.n_colliders
        nb = sc.n_bodies
    if args.tune:
        w.set_param(blobs_b200.abi.PARAM_TUNE, args.tune)
    if args.list is not None:
        w.set_param(blobs_b200.abi.PARAM_LIST, args.list)
    if args.skin is not None:
        w.set_param(blobs_b200.abi.PARAM_SKIN, args.skin)
    if os.environ.get("BLOBS_BENCH_GRAPH", "1") == "0":
        w.set_param(blobs_b200.abi.PARAM_GRAPH, 0)
    if "BLOBS_BENCH_POOL" in os.environ:      # A/B aid: 0 per-lane contact resolution, 1 warp-pooled, 2 auto (library default)
        w.set_param(blobs_b200.abi.PARAM_POOL, int(os.environ["BLOBS_BENCH_POOL"]))
    if "BLOBS_BENCH_POOL_MIN" in os.environ:
        w.set_param(blobs_b200.abi.PARAM_POOL_MIN, int(os.environ["BLOBS_BENCH_POOL_MIN"]))
    if "BLOBS_BENCH_CROWDED" in os.environ:   # A/B aid: 0 inline, 1 always k_crowded, 2 auto (library default)
        w.set_param(blobs_b200.abi.PARAM_CROWDED, int(os.environ["BLOBS_BENCH_CROWDED"]))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local)
    w.step(DT, n=max(W - 3, 0))
    sampler.start()          # nvidia-smi needs ~100 ms to produce its first line: start it under the last warm-up steps (same load)
    w.step(DT, n=min(W, 3))
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    strips_on = world > 1 and args.workload == "cfg2"
    io_cap = (n + n // 8 + 4096) if strips_on else nb     # strips: a rank only exchanges the bodies it owns with its host
    forces = torch.zeros((io_cap, 2), dtype=torch.float32).pin_memory()
    forces[:, 0] = 0.05
    pos_out = torch.zeros((io_cap, 2), dtype=torch.float32).pin_memory()
    slots_io = torch.zeros(io_cap, dtype=torch.int32).pin_memory()

    # ---- device-timed region -----------------------------------------------------------------
    # pass A: K steps, nothing but the step itself between the library's per-step CUDA events (-> value)
    # pass B: the same K steps again with CUDA events around every kernel launch (-> roofline.avg_launch_ms, kernel shares);
    #         the extra event nodes cost a few % of the step, which is why they are kept out of pass A
    bad = [0]

    def timed_pass(profile):
        if profile:
            w.profile_enable(True)
        t_ms, coll, over = 0.0, 0, 0
        for _ in range(K):
            if flush is not None:
                flush.zero_()
                torch.cuda.synchronize()
            st = w.step(DT)
            bad[0] |= st["nan_detected"] & 12  # strip message overflow (4) / peer-exchange wait timed out (8): reported in the JSON line, never raised (a rank that dies would hang its peers)
            t_ms += st["gpu_ms"]
            coll += st["collisions"]
            over += st["list_overflow"]
        return t_ms, coll, over

    barrier()
    sampler.mark()
    nl0 = (w.get_param(blobs_b200.abi.PARAM_LIST_REBUILDS), w.get_param(blobs_b200.abi.PARAM_LIST_SUBSTEPS))
    l0 = w.kernel_info()["launches"]
    t_dev_ms, collisions, overflow = timed_pass(False)
    barrier()
    info = w.kernel_info()
    launches = info["launches"] - l0
    nl1 = (w.get_param(blobs_b200.abi.PARAM_LIST_REBUILDS), w.get_param(blobs_b200.abi.PARAM_LIST_SUBSTEPS))
    t_prof_ms, _, _ = timed_pass(True)
    barrier()
    prof = w.profile_read()
    w.profile_enable(False)
    clocks = sampler.stop()

    # ---- end-to-end region (public C ABI, host buffers, copies inside) ------------------------------
    # N = 1: the pipelined host I/O of the C ABI (blobs_forces_upload_async / blobs_apply_forces_uploaded /
    # blobs_read_body_positions_async / blobs_io_sync): every step's forces are copied from pinned host memory and every step's
    # positions are copied back to pinned host memory inside the timed region, on their own streams, overlapping the kernels of
    # the neighbouring steps. BLOBS_BENCH_E2E=sync times the blocking calls instead (copy, step, copy back to back); a short
    # sample of that loop is always reported beside it as e2e.sync_value.
    io_bytes = 0
    e2e_mode = "sync"
    sync_sample = None
    n_io = w.read_owned_positions_ptr(slots_io.data_ptr(), pos_out.data_ptr(), io_cap) if strips_on else nb

    def sync_loop(k):
        for _ in range(k):
            w.apply_forces_ptr(forces.data_ptr(), nb)      # pinned host -> device
            w.step(DT)
            w.read_positions_ptr(pos_out.data_ptr(), nb)   # device -> pinned host (synchronous)

    def strip_sync_loop(k):
        nonlocal n_io
        moved = 0
        for _ in range(k):
            w.apply_forces_indexed_ptr(slots_io.data_ptr(), forces.data_ptr(), min(n_io, io_cap))        # pinned host -> device (owned bodies)
            moved += min(n_io, io_cap) * 12
            w.step(DT)
            n_io = w.read_owned_positions_ptr(slots_io.data_ptr(), pos_out.data_ptr(), io_cap)            # device -> pinned host (synchronous)
            moved += min(n_io, io_cap) * 12
        return moved

    if strips_on and os.environ.get("BLOBS_BENCH_E2E", "pipelined") == "sync":
        barrier()
        t0 = time.perf_counter()
        io_bytes = strip_sync_loop(K)
        barrier()
        t_e2e = time.perf_counter() - t0
    elif strips_on:
        # N > 1: the same pipeline on the distributed host I/O (blobs_forces_indexed_upload_async / blobs_apply_forces_indexed_uploaded /
        # blobs_read_owned_positions_async): every rank moves the (slot, force) list in and the (slot, position) list out for the
        # bodies it owns, on its own copy streams. The slot list a frame's forces are addressed to is the newest one that has
        # landed on the host (two frames old); a body that migrated in between is skipped by its old owner for that frame.
        ks = max(1, K // 5)
        barrier()
        t0 = time.perf_counter()
        strip_sync_loop(ks)
        barrier()
        sync_sample = (ks, time.perf_counter() - t0)
        e2e_mode = "pipelined"
        sl = (slots_io, slots_io.clone().pin_memory())
        xy = (pos_out, torch.zeros_like(pos_out).pin_memory())
        cnt = (torch.zeros(1, dtype=torch.int32).pin_memory(), torch.zeros(1, dtype=torch.int32).pin_memory())
        cnt[0][0] = cnt[1][0] = min(n_io, io_cap)
        sl[1].copy_(sl[0])
        barrier()
        t0 = time.perf_counter()
        try:
            io_bytes = strip_pipelined_loop(w, K, forces, sl, xy, cnt, io_cap)
            barrier()
            t_e2e = time.perf_counter() - t0
            newest = (K - 1) & 1
            pos_out = xy[newest]
            n_io = int(cnt[newest][0])
        except RuntimeError as e:   # an API-level refusal is the same on every rank: report the blocking loop instead of losing the line
            print(f"bench.py: pipelined strip I/O failed ({e}); e2e falls back to the blocking loop", file=sys.stderr)
            e2e_mode = "sync (pipelined loop refused)"
            barrier()
            t0 = time.perf_counter()
            io_bytes = strip_sync_loop(K)
            barrier()
            t_e2e = time.perf_counter() - t0
    elif os.environ.get("BLOBS_BENCH_E2E", "pipelined") == "sync":
        barrier()
        t0 = time.perf_counter()
        sync_loop(K)
        barrier()
        t_e2e = time.perf_counter() - t0
        io_bytes = K * nb * 16
    else:
        ks = max(1, K // 5)
        barrier()
        t0 = time.perf_counter()
        sync_loop(ks)
        barrier()
        sync_sample = (ks, time.perf_counter() - t0)
        forces2 = forces.clone().pin_memory()
        pos_out2 = torch.zeros_like(pos_out).pin_memory()
        fbuf, obuf = (forces, forces2), (pos_out, pos_out2)
        e2e_mode = "pipelined"
        barrier()
        t0 = time.perf_counter()
        w.forces_upload_async_ptr(fbuf[0].data_ptr(), nb)              # pinned host -> device, step 0
        for i in range(K):
            w.apply_forces_uploaded()
            if i + 1 < K:
                w.forces_upload_async_ptr(fbuf[(i + 1) & 1].data_ptr(), nb)   # step i+1's forces travel under step i's kernels
            w.step(DT)
            w.io_sync()                                                # positions of step i-1 have landed in obuf[(i-1)&1]
            w.read_positions_async_ptr(obuf[i & 1].data_ptr(), nb)     # device -> pinned host, under step i+1's kernels
        w.io_sync()
        barrier()
        t_e2e = time.perf_counter() - t0
        io_bytes = K * nb * 16
        pos_out = obuf[(K - 1) & 1]
    checksum = float(pos_out[: max(1, min(n_io, io_cap)), 1].double().mean())

    t = torch.tensor([t_dev_ms, t_e2e], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(n), float(collisions), float(launches), float(io_bytes), float(bad[0])], dtype=torch.float64, device="cuda")
    mx = torch.tensor([w.get_param(blobs_b200.abi.PARAM_STRIP_MAX_GHOSTS), w.get_param(blobs_b200.abi.PARAM_STRIP_MAX_MIGRANTS)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    t_dev_ms, t_e2e = float(t[0]), float(t[1])
    n_total, coll_total, launches_total = float(tot[0]), float(tot[1]), int(tot[2])

    if rank == 0:
        peak, peak_src = measured_peak()
        main_ms, main_n = prof["main"]
        main_ms += prof["crowded"][0]   # k_crowded finishes the bodies k_main deferred: same algorithmic bytes, so same bucket
        substeps = int(w.get_param(blobs_b200.abi.PARAM_SUBSTEPS))
        achieved = (B_MAIN * n / 1e9) / (main_ms / max(main_n, 1) / 1e3) if main_n else None
        value = n_total * K / (t_dev_ms / 1e3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_dev_ms / K,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc,
                       "spheres_per_gpu": n, "substeps": substeps, "contact_mode": "ordered (bit-exact summation order)",
                       "l2": "256 MiB buffer rewritten between timed steps, outside the per-step CUDA events" if flush is not None else "no flush",
                       "grid": [info["grid_w"], info["grid_h"]], "broadphase_cell": info["broadphase_cell"], "fused_path": info["fused_path"],
                       "contacts_per_step": coll_total / K / max(world, 1), "list_overflow": overflow,
                       "crowded_mode": int(w.get_param(blobs_b200.abi.PARAM_CROWDED)), "pool_mode": int(w.get_param(blobs_b200.abi.PARAM_POOL)),
                       "sim_time_s": [W * DT, (W + K) * DT],
                       "list_mode": int(w.get_param(blobs_b200.abi.PARAM_LIST)), "skin": w.get_param(blobs_b200.abi.PARAM_SKIN),
                       "list_rebuilds_per_substep": (nl1[0] - nl0[0]) / max(nl1[1] - nl0[1], 1.0),
                       "strip_max_ghosts_per_message": int(mx[0]), "strip_max_migrants_per_message": int(mx[1]),
                       "strip_exchange": (("peer-memory stores over NVLink (k_strip_push, CUDA IPC)" if int(w.get_param(blobs_b200.abi.PARAM_STRIP_P2P)) else "grouped ncclSend/ncclRecv")
                                          if strips_on else None),
                       "main_kernel": VARIANT_NAME.get(args.tune, f"k_main (tune {args.tune})"), "autotune": tune_report},
            "clocks": clocks,
            "e2e": {"value": n_total * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": float(tot[3]) / K / 2, "d2h_bytes_per_step": float(tot[3]) / K / 2, "ms_per_step": t_e2e / K * 1e3,
                    "checksum_mean_y": checksum, "host_io": e2e_mode,
                    "sync_value": (n_total * sync_sample[0] / sync_sample[1]) if sync_sample else None},
            "gpu_launches": launches_total,
            "roofline": {"bound": "hbm", "kernel": ("k_tile" if args.tune in (11, 12) else "k_main<fused,ordered>") + " (contacts + verlet + snapshot + clamp + cell binning)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": B_MAIN * n, "avg_launch_ms": main_ms / max(main_n, 1),
                         "timing": "CUDA events around every k_main launch, second timed pass of the same K steps",
                         "traffic": ncu_traffic() if args.tune in (0, 2, 3, 4, 5, 6, 7) else None},   # the committed ncu capture is k_main's (per-lane variant)
            "pipeline": {"algorithmic_gbps": B_PIPELINE * n_total * substeps * K / (t_dev_ms / 1e3) / 1e9,
                         "frac_of_peak": B_PIPELINE * n_total * substeps * K / (t_dev_ms / 1e3) / 1e9 / (peak * world),
                         "sphere_substeps_per_sec": value * substeps,
                         "kernel_ms_per_step": {k: v[0] / K for k, v in prof.items() if v[1]},
                         "ms_per_step_with_kernel_events": t_prof_ms / K, "cuda_graph_replays": int(w.get_param(blobs_b200.abi.PARAM_GRAPH_REPLAYS))},
        }
        if float(tot[4]) != 0:
            line["invalid"] = "strip message buffers overflowed (raise ghost_capacity / migrate_capacity) or a peer-memory exchange timed out: results are not valid"
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_sample(args.cpu_budget)[0]
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    try:
        if args.probe_strips:
            run_probe_strips(args)
        elif args.probe:
            run_probe(args)
        elif args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    except BaseException:  # noqa: BLE001 — under torchrun a rank must die at once, or its peers wait in NCCL for ever
        import traceback

        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)
    sys.stdout.flush()
    os._exit(0)   # skip interpreter teardown: nothing may block on a stream that sits in a collective


if __name__ == "__main__":
    main()
