#!/usr/bin/env python
"""bench.py — sphere-steps/sec of the blobs::Physics::step hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg2|cfg3|cfg4|...]

A "step" is one Physics::step(1/60) (8 substeps) over a synthetic scene. N = 1: BASELINE config #2 (1 048 576 single-collider
spheres in one world). N > 1: the config #5 family (ONE world of 2 097 152 spheres per GPU, strip-decomposed, ghost / migration
exchange every substep); `--workload cfg3` = 4096 batched independent worlds partitioned over the ranks. Prints ONE JSON line.

`value`   device time of K steps (CUDA events recorded by the library on its own stream around each step), inputs resident in
          HBM, max over ranks; a 256 MiB buffer is rewritten between timed steps (L2 flush, outside the events).
`e2e`     wall clock through the public C ABI with, every step, forces copied from pinned host memory and positions copied
          back to pinned host memory inside the timed region (the ABI's pipelined host I/O; `sync_value` = the blocking calls).
`roofline` the dominant kernel (k_step or k_main, whichever pipeline the library chose): 116 algorithmic bytes per collider
          and launch over its average launch time, from live CUDA events in a second pass over the same K steps.
`config.steady_state` / `config.late_state`: the same device-timed measurement at the SURVEY window (steps 200-300 of the
          simulation) and at steps 700-800 — the scene is not stationary: the block falls freely for ~110 steps, hits the circle
          constraint and is turned into a hot gas by the reference's positional solver. `value` is whatever window --warmup /
          --steps select; these two say what the same code does later in the same simulation.
`cpu_baseline` the reference's algorithm (O(C^2) pair loop, one thread: the reference is !Send) as restated by the CPU oracle, on
          a bounded sample; `cpu_baseline.grid_restatement` = all host cores, cell lists (NOT the reference algorithm), full size.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "sphere_steps_per_sec"
UNIT = "sphere-steps/s"
DT = 1.0 / 60.0
B_MAIN = 64 + 52      # algorithmic bytes per collider-substep of the fused contact + verlet + snapshot + clamp + key kernel
B_PIPELINE = 236      # ... of the whole substep


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg2_varied", "cfg2_dense", "cfg3", "cfg4", "cfg2_4m", "cfg2_16m"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--no-late", action="store_true", help="skip the steady_state / late_state windows")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the strip-world-vs-one-GPU self-check (parity_vs_single_gpu)")
    ap.add_argument("--tune", type=int, default=0, help="kernel variant selector (BLOBS_PARAM_TUNE)")
    ap.add_argument("--list", type=int, default=None, help="BLOBS_PARAM_LIST: 0 = cell grid every substep, 1 = neighbour lists, 2 = automatic (library default)")
    ap.add_argument("--skin", type=float, default=None, help="BLOBS_PARAM_SKIN (fraction of the largest radius)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the cpu_baseline sample")
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
        self.n_before = 0

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
                "samples_before_timed_region": self.n_before, "interval_ms": 20, "power_w_max": max(power)}


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, n_colliders, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from a committed `ncu --set full` capture of the SAME kernel on the
    SAME workload at the SAME size (profiles/kernel_traffic.json), else None."""
    try:
        for e in json.load(open(os.path.join(REPO, "profiles", "kernel_traffic.json")))["captures"]:
            if e["kernel"] == kernel and int(e["colliders"]) == int(n_colliders) and e.get("workload") == workload:
                return float(e["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def cpu_reference_sample(budget_s, steps_hint=None, warmup=0):
    """The reference's own algorithm (brute-force O(C^2) Physics::step; CPU oracle = faithful restatement; 1 thread, the reference
    is !Send) on a bounded sample of cfg2: a side x side sub-block of the same lattice."""
    from blobs_b200 import scenes as S
    from oracle import oracle_py

    def run(side, steps, wu):
        sc = S.cfg2(seed=1, side=side)
        o = oracle_py.OracleWorld(gravity=sc.gravity, maintain_spatial_hash=True, record_events=True)
        S.build(o, sc)
        if wu:
            o.step_n_timed(DT, wu)
            o.events_drain(); o.pairs_drain()
        return sc.n_colliders, o.step_n_timed(DT, steps)

    n0, s0 = run(32, 2, 0)   # calibrate the pair-test rate on a tiny block
    rate = (n0 * n0 / 2.0) * 8 * 2 / max(s0, 1e-6)
    if steps_hint is None:
        side = 96
        steps = max(1, min(50, int(budget_s / max((side * side) ** 2 / 2.0 * 8 / rate, 1e-6))))
    else:
        steps = steps_hint
        per_step_budget = budget_s / max(1, steps + warmup)
        side = int(max(16, min(128, ((2.0 * per_step_budget * rate / 8.0) ** 0.5) ** 0.5)))
    n, secs = run(side, steps, warmup)
    full = 1048576
    return {
        "value": n * steps / secs, "unit": UNIT, "cores": 1, "kind": "port",
        "sample": f"{n} spheres ({side}x{side} sub-block of the cfg2 lattice) x {steps} Physics::step(1/60), brute-force O(C^2) pair loop as in the "
                  f"reference (physics.rs:241-317), single thread, {secs:.2f} s; host has {os.cpu_count()} cores; the reference cannot be built here "
                  f"(Rust), so this is the C++ oracle restatement. O(C^2): at the full {full} spheres the same loop would run ~{full / n:.0f}x slower per sphere",
        "pair_tests_per_s": rate,
    }, n, steps, secs


def cpu_grid_restatement(steps=2):
    """All host cores, cell lists: oracle/grid_omp.cpp on the FULL cfg2 scene. Labelled: not the reference algorithm (same
    arithmetic and summation order, bit-identical results; tests/test_grid_omp.py)."""
    from blobs_b200 import scenes as S
    from oracle import grid_omp

    sc = S.cfg2(seed=1)
    g = grid_omp.GridOmpWorld(sc)
    g.step(DT, n=1)   # first-touch / warm caches
    r = g.step(DT, n=steps)
    return {"value": sc.n_bodies * steps / r["seconds"], "unit": UNIT, "cores": grid_omp.max_threads(), "kind": "port",
            "sample": f"CPU cell-list restatement (OpenMP, oracle/grid_omp.cpp; NOT the reference algorithm, which is O(C^2) on one thread): the full "
                      f"{sc.n_bodies}-sphere cfg2 scene x {steps} Physics::step(1/60) after 1 warm-up step, {r['seconds']:.2f} s"}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cb, n, steps, secs = cpu_reference_sample(120.0, steps_hint=args.steps, warmup=args.warmup)
    try:
        cb["grid_restatement"] = cpu_grid_restatement()
    except Exception as e:  # noqa: BLE001 - context only
        cb["grid_restatement"] = {"unavailable": str(e)[:200]}
    v = cb["value"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2 (bounded sample): " + cb["sample"]},
        "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
    }), flush=True)


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


def strip_parity_probe(dist, rank, world, local, dev, steps=20, ny=192, cols_per_rank=48, params=None):
    """Parity at N > 1, visible in the bench line: a small strip world (cols_per_rank*world x ny spheres with lateral velocities, so
    bodies migrate across the edges) stepped on all ranks with the same settings as the timed world, against the same world on rank
    0's GPU alone; positions, old positions and velocities of every body compared bit for bit. (The test of record is
    tests/test_multi_gpu.py; this is the same comparison, run by the bench itself.)"""
    import numpy as np
    import torch

    import blobs_b200
    from blobs_b200 import scenes as S
    from blobs_b200 import strips

    A = blobs_b200.abi
    nx = cols_per_rank * world
    sc = S.lattice_scene(nx, ny, 1.05, (0.0, 0.0), 5, 0.3, 0.5, jitter=0.04, vel_disc=6.0, constraint_r=0.8 * max(nx, ny), name="strip-probe", cell_size=1.0)

    def make():
        w = blobs_b200.World(gravity=sc.gravity, device=local)
        S.build(w, sc)
        for k, v in (params or {}).items():
            w.set_param(k, v)
        return w

    def state(w):
        b, _ = w.download_bodies()
        return np.concatenate([b["position"]["x"], b["position"]["y"], b["position_old"]["x"], b["position_old"]["y"],
                               b["calculated_velocity"]["x"], b["calculated_velocity"]["y"]]).astype(np.float32)

    w = make()
    edges = strips.strip_edges(float(sc.bodies["position"]["x"].min()), float(sc.bodies["position"]["x"].max()), world)
    uid = torch.from_numpy(blobs_b200.World.strip_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).to(dev)
    dist.broadcast(uid, 0)
    w.strip_configure(rank, world, float(edges[rank]), float(edges[rank + 1]), uid.cpu().numpy(), ghost_capacity=1 << 14, migrate_capacity=1 << 10)
    own0 = w.strip_owned().astype(bool)
    bad = coll = 0
    for _ in range(steps):
        st = w.step(DT)
        bad |= st["nan_detected"]
        coll += st["collisions"]
    own = w.strip_owned().astype(bool)
    lists = int(w.get_param(A.PARAM_LIST_ACTIVE))
    mine = np.where(np.tile(own, 6), state(w), np.float32(0)).view(np.int32)   # a body's fields are non-zero on its owner only
    t = torch.from_numpy(mine.astype(np.int64)).to(dev)
    o = torch.from_numpy(np.concatenate([own, own0 != own]).astype(np.int64)).to(dev)
    c = torch.tensor([coll, bad], dtype=torch.int64, device=dev)
    dist.all_reduce(t)      # exact: every slot has one non-zero contributor (int64 sums of int32 bit patterns)
    dist.all_reduce(o)
    dist.all_reduce(c)
    w.close()               # only now: a neighbour's last substep may still have been storing ghost records into this rank's buffers
    out = None
    if rank == 0:
        n = sc.n_bodies
        ref = make()
        ref_coll = 0
        for _ in range(steps):
            ref_coll += ref.step(DT)["collisions"]
        want = state(ref).view(np.int32).astype(np.int64)
        ref.close()
        owners = o.cpu().numpy()
        mism = int((t.cpu().numpy() != want).sum())
        out = {"spheres": n, "steps": steps, "ranks": world, "list_pipeline_active": lists, "migrated_bodies": int(owners[n:].sum()) // 2,
               "contacts": int(c[0]), "contacts_single_gpu": ref_coll, "mismatching_words": mism,
               "bit_identical": bool(mism == 0 and int(c[0]) == ref_coll and int(c[1]) == 0 and (owners[:n] == 1).all())}
    return out


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

    A = blobs_b200.abi
    scaling = "weak"
    strips_on = world > 1 and args.workload == "cfg2"
    if args.workload == "cfg3":
        # BASELINE config #3: 4096 independent worlds x 256 bodies, block-partitioned over the ranks, no collective (strong scaling)
        n_worlds_total = 4096
        lo, hi = rank * n_worlds_total // world, (rank + 1) * n_worlds_total // world
        worlds = [S.cfg1(1 + wid, n_side=16) for wid in range(lo, hi)]
        def make_world():
            ww = blobs_b200.World(gravity=worlds[0].gravity, device=local, body_capacity=256 * (hi - lo), collider_capacity=256 * (hi - lo))
            S.build_batch(ww, worlds)
            return ww

        w = make_world()
        n = nb = 256 * (hi - lo)
        desc = f"cfg3: {n_worlds_total} batched independent worlds x 256 bodies (cfg1 at 16x16, circle R=4), worlds {lo}..{hi - 1} on this rank"
        scaling = "strong"
    elif strips_on:
        # BASELINE config #5 family: ONE world of 512*N x 4096 spheres (N=8: 16 777 216), cut into N vertical strips of 512 lattice
        # columns (2 097 152 spheres per GPU, weak scaling); ghost + migration exchange with both neighbours every substep.
        from blobs_b200 import strips

        nx, ny = 512 * world, 4096
        sc = S.lattice_scene(nx, ny, 1.05, (0.0, 0.0), 1, 0.5, 0.5, jitter=0.04, vel_disc=1.0, constraint_r=0.8 * max(nx, ny), name="cfg5", cell_size=1.0)
        if os.environ.get("BLOBS_BENCH_STRIP_MAJOR", "1") != "0":
            # Insert the bodies strip by strip (column block of 512, then row, then column) instead of row by row: every array of a
            # strip world is indexed by GLOBAL slot on every rank, so this makes a rank's share of each array one contiguous range
            # instead of 4 KB out of every 32 KB - 8x fewer pages touched per rank at N = 8. Same bodies, same attributes (the
            # scene's RNG is keyed by the lattice index); only the slot numbering - the order of the reference's pair loop - differs.
            idx = np.arange(nx * ny)
            ix, iy = idx % nx, idx // nx
            perm = np.lexsort((ix % 512, iy, ix // 512))
            sc.bodies, sc.colliders = sc.bodies[perm], sc.colliders[perm]
            del idx, ix, iy, perm
        desc = (f"cfg5 family: one world of {nx}x{ny} = {nx * ny} spheres r=0.5 (pitch 1.05, circle R={0.8 * max(nx, ny):.0f}), strip-decomposed over {world} GPUs "
                f"({nx * ny // world} spheres per GPU), bodies inserted strip-major, ghost/migration exchange with both neighbours every substep")
        w = blobs_b200.World(gravity=sc.gravity, device=local, body_capacity=sc.n_bodies, collider_capacity=sc.n_colliders)
        S.build(w, sc)
        if os.environ.get("BLOBS_B200_STRIP_P2P", "1") != "0":
            w.set_param(A.PARAM_STRIP_P2P, 1)   # peer-memory exchange (falls back to NCCL on every rank if any rank cannot map its neighbours)
        edges = strips.strip_edges(float(sc.bodies["position"]["x"].min()), float(sc.bodies["position"]["x"].max()), world)
        uid = torch.from_numpy(blobs_b200.World.strip_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).cuda()
        dist.broadcast(uid, 0)
        w.strip_configure(rank, world, float(edges[rank]), float(edges[rank + 1]), uid.cpu().numpy(), ghost_capacity=4 * ny, migrate_capacity=2 * ny)
        n = int(w.strip_owned().sum())
        nb = sc.n_bodies
        del sc
    else:
        # other workloads at N > 1: one independent world per rank (no data-path collective; "replicas", weak scaling)
        sc, desc = make_scene(args.workload, seed=1 + rank)
        if world > 1:
            desc += f"; one independent world per GPU ({world} replicas, no collective)"
        def make_world():
            ww = blobs_b200.World(gravity=sc.gravity, device=local, body_capacity=sc.n_bodies, collider_capacity=sc.n_colliders)
            S.build(ww, sc)
            return ww

        w = make_world()
        n, nb = sc.n_colliders, sc.n_bodies
    knobs = {}
    if args.tune:
        knobs[A.PARAM_TUNE] = args.tune
    if args.list is not None:
        knobs[A.PARAM_LIST] = args.list
    if args.skin is not None:
        knobs[A.PARAM_SKIN] = args.skin
    for k, v in knobs.items():
        w.set_param(k, v)
    if os.environ.get("BLOBS_BENCH_GRAPH", "1") == "0":
        w.set_param(A.PARAM_GRAPH, 0)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local) if rank == 0 else None   # one nvidia-smi poller per node: eight of them stall each other's CUDA calls
    w.step(DT, n=max(W - 3, 0))
    if sampler:
        sampler.start()      # nvidia-smi needs ~100 ms to produce its first line: start it under the last warm-up steps (same load)
    w.step(DT, n=min(W, 3))
    sim_steps = W
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    io_cap = (n + n // 8 + 4096) if strips_on else nb     # strips: a rank only exchanges the bodies it owns with its host
    forces = torch.zeros((io_cap, 2), dtype=torch.float32).pin_memory()
    forces[:, 0] = 0.05
    pos_out = torch.zeros((io_cap, 2), dtype=torch.float32).pin_memory()
    slots_io = torch.zeros(io_cap, dtype=torch.int32).pin_memory()
    bad = [0]

    def timed_pass(k, profile=False):
        """k steps, nothing but the step itself between the library's per-step CUDA events; `profile` adds CUDA events around every
        kernel launch (they cost a few % of the step, which is why `value` comes from a pass without them)."""
        if profile:
            w.profile_enable(profile)   # 1 / True: every kernel class, 2: the dominant kernel only
        r0 = (w.get_param(A.PARAM_LIST_REBUILDS), w.get_param(A.PARAM_LIST_SUBSTEPS))
        t_ms, coll, over = 0.0, 0, 0
        for _ in range(k):
            if flush is not None:
                flush.zero_()
                torch.cuda.synchronize()
            st = w.step(DT)
            bad[0] |= st["nan_detected"] & 12  # strip message overflow (4) / peer-exchange wait timed out (8): reported, never raised (a rank that dies would hang its peers)
            t_ms += st["gpu_ms"]
            coll += st["collisions"]
            over += st["list_overflow"]
        prof = None
        if profile:
            prof = w.profile_read()
            w.profile_enable(False)
        r1 = (w.get_param(A.PARAM_LIST_REBUILDS), w.get_param(A.PARAM_LIST_SUBSTEPS))
        return {"ms": t_ms, "collisions": coll, "overflow": over, "prof": prof, "list_active": int(w.get_param(A.PARAM_LIST_ACTIVE)),
                "rebuilds_per_substep": (r1[0] - r0[0]) / max(r1[1] - r0[1], 1.0)}

    # ---- device-timed region ----------------------------------------------------------------------------------------
    barrier()
    if sampler:
        sampler.mark()
    l0 = w.kernel_info()["launches"]
    if strips_on:
        # Strip worlds: the K steps are enqueued by ONE call (blobs_step_n) and timed by its pair of CUDA events. No L2 flush is needed
        # (2 M spheres per GPU: ~270 MB touched per substep against 126 MB of L2), and without a host round trip per step the ranks
        # do not pick up each other's host-side launch jitter: the list pipeline synchronises ALL ranks at the start of every substep,
        # so a rank that enters a step late would stall the seven others inside their timed region.
        r0 = (w.get_param(A.PARAM_LIST_REBUILDS), w.get_param(A.PARAM_LIST_SUBSTEPS))
        st = w.step(DT, n=K)
        bad[0] |= st["nan_detected"] & 12
        r1 = (w.get_param(A.PARAM_LIST_REBUILDS), w.get_param(A.PARAM_LIST_SUBSTEPS))
        main = {"ms": st["gpu_ms"], "collisions": st["collisions"], "overflow": st["list_overflow"], "prof": None,
                "list_active": int(w.get_param(A.PARAM_LIST_ACTIVE)), "rebuilds_per_substep": (r1[0] - r0[0]) / max(r1[1] - r0[1], 1.0)}
    else:
        main = timed_pass(K)
    barrier()
    info = w.kernel_info()
    launches = info["launches"] - l0
    profd = timed_pass(K, profile=True)
    barrier()
    # events around the dominant kernel only, plain launches (inside a replayed graph the event-record nodes add several us per pair)
    w.set_param(A.PARAM_GRAPH, 0)
    profm = timed_pass(K, profile=2)
    if os.environ.get("BLOBS_BENCH_GRAPH", "1") != "0":
        w.set_param(A.PARAM_GRAPH, 1)
    barrier()
    clocks = sampler.stop() if sampler else None
    sim_steps += 3 * K

    # ---- end-to-end region (public C ABI, host buffers, copies inside) --------------------------------------------------
    # A plain world gets a fresh copy of the scene for this region, warmed up like the first, so that e2e is measured over the same
    # stretch of the simulation as `value` (the scene is not stationary; the passes above have advanced `w` by 3 K steps). A strip
    # world carries on (its set-up is collective).
    io_bytes = 0
    w_main = w
    if not strips_on:
        w = make_world()
        for k, v in knobs.items():
            w.set_param(k, v)
        if os.environ.get("BLOBS_BENCH_GRAPH", "1") == "0":
            w.set_param(A.PARAM_GRAPH, 0)
        w.step(DT, n=W)
    n_io = w.read_owned_positions_ptr(slots_io.data_ptr(), pos_out.data_ptr(), io_cap) if strips_on else nb

    def sync_loop(k):
        nonlocal n_io
        moved = 0
        for _ in range(k):
            if strips_on:
                w.apply_forces_indexed_ptr(slots_io.data_ptr(), forces.data_ptr(), min(n_io, io_cap))   # pinned host -> device (owned bodies)
                moved += min(n_io, io_cap) * 12
                w.step(DT)
                n_io = w.read_owned_positions_ptr(slots_io.data_ptr(), pos_out.data_ptr(), io_cap)        # device -> pinned host (synchronous)
                moved += min(n_io, io_cap) * 12
            else:
                w.apply_forces_ptr(forces.data_ptr(), nb)
                w.step(DT)
                w.read_positions_ptr(pos_out.data_ptr(), nb)
                moved += nb * 16
        return moved

    ks = max(1, K // 5)
    barrier()
    t0 = time.perf_counter()
    sync_loop(ks)
    barrier()
    sync_sample = (ks, time.perf_counter() - t0)
    e2e_mode = "pipelined"
    barrier()
    # The timed window is K steps, like `value`; it is preceded by 2 untimed steps of the same loop (the pipelined entry points create
    # their copy streams and staging buffers on first use) and repeated 3 times: 20 steps are 10-30 ms of wall clock, and one host
    # hiccup in them moved e2e by 2x between otherwise identical runs. e2e = the MEDIAN window; all three are reported.
    E2E_WINDOWS = 3
    if strips_on:
        # every rank moves the (slot, force) list in and the (slot, position) list out for the bodies it owns, on its own copy
        # streams; a frame's forces are addressed to the newest slot list that has landed on the host (two frames old)
        sl = (slots_io, slots_io.clone().pin_memory())
        xy = (pos_out, torch.zeros_like(pos_out).pin_memory())
        cnt = (torch.zeros(1, dtype=torch.int32).pin_memory(), torch.zeros(1, dtype=torch.int32).pin_memory())
        cnt[0][0] = cnt[1][0] = min(n_io, io_cap)

        def e2e_window(k):
            return strip_pipelined_loop(w, k, forces, sl, xy, cnt, io_cap)
    else:
        fbuf, obuf = (forces, forces.clone().pin_memory()), (pos_out, torch.zeros_like(pos_out).pin_memory())

        def e2e_window(k):
            w.forces_upload_async_ptr(fbuf[0].data_ptr(), nb)              # pinned host -> device, step 0
            for i in range(k):
                w.apply_forces_uploaded()
                if i + 1 < k:
                    w.forces_upload_async_ptr(fbuf[(i + 1) & 1].data_ptr(), nb)   # step i+1's forces travel under step i's kernels
                w.step(DT)
                w.io_sync()                                                # positions of step i-1 have landed in obuf[(i-1)&1]
                w.read_positions_async_ptr(obuf[i & 1].data_ptr(), nb)     # device -> pinned host, under step i+1's kernels
            w.io_sync()
            return k * nb * 16

    e2e_window(2)
    e2e_times = []
    for _ in range(E2E_WINDOWS):
        barrier()
        t0 = time.perf_counter()
        io_bytes = e2e_window(K)
        barrier()
        e2e_times.append(time.perf_counter() - t0)
    if strips_on:
        pos_out, n_io = xy[(K - 1) & 1], int(cnt[(K - 1) & 1][0])
    else:
        pos_out = obuf[(K - 1) & 1]
    if strips_on:
        sim_steps += ks + 2 + E2E_WINDOWS * K
    else:
        w.close()
        w = w_main
    checksum = float(pos_out[: max(1, min(n_io, io_cap)), 1].double().mean())

    # ---- later windows of the same simulation (single world only: a strip world would need its collective bookkeeping) -------
    windows = {}
    if not args.no_late and world == 1 and args.workload.startswith("cfg2"):
        for name, start in (("steady_state", 200), ("late_state", 700)):
            if sim_steps > start:
                continue
            w.step(DT, n=start - sim_steps)
            r = timed_pass(100)
            p = timed_pass(20, profile=True)
            w.set_param(A.PARAM_GRAPH, 0)
            pm = timed_pass(20, profile=2)
            if os.environ.get("BLOBS_BENCH_GRAPH", "1") != "0":
                w.set_param(A.PARAM_GRAPH, 1)
            sim_steps = start + 140
            mm, mn = pm["prof"]["main"]
            mm += pm["prof"]["crowded"][0]
            windows[name] = {"value": n * 100 / (r["ms"] / 1e3), "unit": UNIT, "ms_per_step": r["ms"] / 100, "sim_steps": [start, start + 100],
                             "contacts_per_step": r["collisions"] / 100, "list_pipeline_active": r["list_active"],
                             "list_rebuilds_per_substep": r["rebuilds_per_substep"],
                             "main_kernel_avg_launch_ms": mm / max(mn, 1), "roofline_frac": (B_MAIN * n / 1e9) / (mm / max(mn, 1) / 1e3) / measured_peak()[0] if mn else None,
                             "kernel_ms_per_step": {k: v[0] / 20 for k, v in p["prof"].items() if v[1]}}

    # ---- the N > 1 family's per-GPU problem on ONE GPU (2 097 152 spheres, same lattice and circle as one strip of the 8-GPU world):
    # the denominator a weak-scaling efficiency needs (the N = 1 line itself is config #2, half that size) ----------------------
    same_size = None
    if not args.no_late and world == 1 and args.workload == "cfg2":
        sc2 = S.lattice_scene(512, 4096, 1.05, (0.0, 0.0), 1, 0.5, 0.5, jitter=0.04, vel_disc=1.0, constraint_r=0.8 * 4096, name="cfg5-one-strip", cell_size=1.0)
        w2 = blobs_b200.World(gravity=sc2.gravity, device=local, body_capacity=sc2.n_bodies, collider_capacity=sc2.n_colliders)
        S.build(w2, sc2)
        if args.list is not None:
            w2.set_param(A.PARAM_LIST, args.list)
        w2.step(DT, n=W)
        ms2 = 0.0
        for _ in range(K):
            if flush is not None:
                flush.zero_()
                torch.cuda.synchronize()
            ms2 += w2.step(DT)["gpu_ms"]
        same_size = {"value": sc2.n_bodies * K / (ms2 / 1e3), "unit": UNIT, "ms_per_step": ms2 / K, "spheres": sc2.n_bodies, "sim_steps": [W, W + K],
                     "what": "one GPU, 512 x 4096 lattice (the per-GPU share of the N > 1 strip worlds), same timing method as `value`"}
        w2.close()
        del w2, sc2

    # ---- N > 1: the strip decomposition against one GPU, same settings, small world (bit-exact or the line says so) ------------
    parity = None
    if strips_on and not args.no_parity:
        if os.environ.get("BLOBS_B200_STRIP_P2P", "1") != "0":
            knobs[A.PARAM_STRIP_P2P] = 1
        parity = strip_parity_probe(dist, rank, world, local, "cuda", params=knobs)

    t = torch.tensor([main["ms"]] + e2e_times, dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(n), float(main["collisions"]), float(launches), float(io_bytes), float(bad[0])], dtype=torch.float64, device="cuda")
    mx = torch.tensor([w.get_param(A.PARAM_STRIP_MAX_GHOSTS), w.get_param(A.PARAM_STRIP_MAX_MIGRANTS)], dtype=torch.float64, device="cuda")
    t_min = t.clone()
    if dist is not None:
        dist.all_reduce(t_min, op=dist.ReduceOp.MIN)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    t_dev_ms = float(t[0])
    e2e_windows = sorted(float(x) for x in t[1:])     # each window: max over ranks
    t_e2e = e2e_windows[len(e2e_windows) // 2]
    n_total, coll_total, launches_total = float(tot[0]), float(tot[1]), int(tot[2])

    if rank == 0:
        peak, peak_src = measured_peak()
        prof = profd["prof"]
        main_ms, main_n = profm["prof"]["main"]
        main_ms += profm["prof"]["crowded"][0]   # k_crowded finishes the bodies the main kernel deferred: same algorithmic bytes, same bucket
        substeps = int(w.get_param(A.PARAM_SUBSTEPS))
        kernel = "k_step" if main["list_active"] else "k_main"
        achieved = (B_MAIN * n / 1e9) / (main_ms / max(main_n, 1) / 1e3) if main_n else None
        value = n_total * K / (t_dev_ms / 1e3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_dev_ms / K,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "spheres_per_gpu": n, "substeps": substeps, "contact_mode": "ordered (bit-exact summation order)",
                       "l2": ("no flush: K steps enqueued by one blobs_step_n call; the per-GPU working set (2 M spheres, ~270 MB per substep) exceeds the 126 MB L2" if strips_on
                              else "256 MiB buffer rewritten between timed steps, outside the per-step CUDA events" if flush is not None else "no flush"),
                       "grid": [info["grid_w"], info["grid_h"]], "broadphase_cell": info["broadphase_cell"], "fused_path": info["fused_path"],
                       "sim_steps": [W, W + K], "ms_per_step_fastest_rank": float(t_min[0]) / K, "contacts_per_step": coll_total / K / max(world, 1), "list_overflow": main["overflow"],
                       "broadphase": ("neighbour lists (k_step), %.3f rebuilds per substep" % main["rebuilds_per_substep"]) if main["list_active"]
                                     else "cell grid rebuilt every substep (k_main)",
                       "list_mode": int(w.get_param(A.PARAM_LIST)), "skin": w.get_param(A.PARAM_SKIN),
                       "strip_max_ghosts_per_message": int(mx[0]), "strip_max_migrants_per_message": int(mx[1]),
                       "strip_exchange": (("peer-memory stores over NVLink (k_strip_push, CUDA IPC)" if int(w.get_param(A.PARAM_STRIP_P2P)) else "grouped ncclSend/ncclRecv")
                                          if strips_on else None),
                       "kernel_ms_per_step": {k: v[0] / K for k, v in prof.items() if v[1]},
                       "pipeline_algorithmic_gbps": B_PIPELINE * n_total * substeps * K / (t_dev_ms / 1e3) / 1e9,
                       "pipeline_frac_of_peak": B_PIPELINE * n_total * substeps * K / (t_dev_ms / 1e3) / 1e9 / (peak * world),
                       "ms_per_step_with_kernel_events": profd["ms"] / K, "cuda_graph_replays": int(w.get_param(A.PARAM_GRAPH_REPLAYS)),
                       **windows, **({"one_gpu_at_strip_size": same_size} if same_size else {})},
            "clocks": clocks,
            "e2e": {"value": n_total * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": float(tot[3]) / K / 2, "d2h_bytes_per_step": float(tot[3]) / K / 2,
                    "ms_per_step": t_e2e / K * 1e3, "windows_ms_per_step": [x / K * 1e3 for x in e2e_windows], "window": "median of 3 consecutive windows of K steps, after 2 untimed steps of the same loop" + ("" if strips_on else "; a fresh copy of the scene warmed up like the first, simulation steps %d..%d" % (W + ks + 2, W + ks + 2 + 3 * K)),
                    "checksum_mean_y": checksum, "host_io": e2e_mode, "sync_value": n_total * sync_sample[0] / sync_sample[1]},
            "gpu_launches": launches_total,
            "roofline": {"bound": "hbm", "kernel": kernel + " (contacts + verlet + snapshot + clamp" + (")" if main["list_active"] else " + cell binning)"),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": B_MAIN * n, "avg_launch_ms": main_ms / max(main_n, 1),
                         "timing": "CUDA events around every launch of this kernel (and no other), plain launches, a third pass over K steps of the same window",
                         "traffic": ncu_traffic(kernel, n, args.workload)},
        }
        if parity is not None:
            line["parity_vs_single_gpu"] = parity
            if not parity["bit_identical"]:
                line["invalid"] = "the strip world differs from the same world on one GPU (parity_vs_single_gpu)"
        if float(tot[4]) != 0:
            line["invalid"] = "strip message buffers overflowed (raise ghost_capacity / migrate_capacity) or a peer-memory exchange timed out: results are not valid"
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_sample(args.cpu_budget)[0]
            try:
                line["cpu_baseline"]["grid_restatement"] = cpu_grid_restatement()
            except Exception as e:  # noqa: BLE001 - context only
                line["cpu_baseline"]["grid_restatement"] = {"unavailable": str(e)[:200]}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    w.close()


def main():
    args = parse()
    try:
        if args.impl == "reference":
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


if __name__ == "__main__":
    main()
