"""Worker for the multi-GPU parity test (run under torchrun, one rank per GPU):
a strip-decomposed world across all ranks must be bit-identical to the same world on one GPU.
With BLOBS_TEST_EMU=1 the same worker runs WITHOUT GPUs: every rank is a CPU process backed by the host-compiled build of the
CUDA sources (tests/emu), and the per-substep ghost / migration exchange goes through a socket stand-in for NCCL."""
import contextlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    emu = os.environ.get("BLOBS_TEST_EMU") == "1"
    if emu:
        sys.path.insert(0, os.path.join(REPO, "tests"))
        import emu_loader

        os.environ["BLOBS_EMU_NCCL_LIB"] = os.path.join(emu_loader.EMU_DIR, "libnccl_fake.so")
        backend = emu_loader.emulated()
    else:
        torch.cuda.set_device(local)
        backend = contextlib.nullcontext()
    dist.init_process_group("gloo")
    with backend:
        return run(rank, world, local)


def run(rank, world, local):
    import blobs_b200
    from blobs_b200 import _abi as A
    from blobs_b200 import scenes as S
    from blobs_b200 import strips

    side = int(os.environ.get("STRIP_TEST_SIDE", "192"))
    steps = int(os.environ.get("STRIP_TEST_STEPS", "40"))
    # gas of spheres with lateral velocities (so bodies really migrate across strip edges) inside a roomy circle
    if os.environ.get("STRIP_TEST_SCENE") == "shell":
        # the block is larger than its circle: the clamp builds a dense shell on the boundary, cut by the strip edge, so bodies
        # with more contacts than the in-register list holds (k_crowded, incl. its ghost / migrant packing) sit on both sides
        sc = S.lattice_scene(side, side, 1.05, (0.0, 0.0), 5, 0.5, 0.5, jitter=0.04, vel_disc=6.0, constraint_r=0.42 * side, name="strip-shell", cell_size=1.0)
    else:
        sc = S.lattice_scene(side, side, 1.05, (0.0, 0.0), 5, 0.3, 0.5, jitter=0.04, vel_disc=6.0, constraint_r=0.8 * side, name="strip", cell_size=1.0)
    w = blobs_b200.World(gravity=sc.gravity, device=local)
    S.build(w, sc)
    edges = strips.agree_edges(dist, sc.bodies["position"]["x"], world)
    assert strips.check_strip_width(edges, 0.5)
    uid = torch.from_numpy(blobs_b200.World.strip_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8))
    dist.broadcast(uid, 0)
    w.strip_configure(rank, world, float(edges[rank]), float(edges[rank + 1]), uid.numpy(), ghost_capacity=1 << 14, migrate_capacity=1 << 10)
    own0 = w.strip_owned()
    try:   # events on a strip world are refused in either order (strip_configure refuses a world that already records events)
        w.record_contacts(A.RECORD_EVENTS, 16)
        raise AssertionError("record_contacts(EVENTS) must be refused on a strip world")
    except blobs_b200.BlobsError:
        pass
    p2p = int(w.get_param(A.PARAM_STRIP_P2P))   # 1 = ghosts / migrants travel by peer-memory stores (k_strip_push), 0 = ncclSend/ncclRecv
    if os.environ.get("STRIP_TEST_EXPECT_P2P"):
        assert p2p == int(os.environ["STRIP_TEST_EXPECT_P2P"]), f"rank {rank}: peer-memory exchange active={p2p}"
    lists = int(w.get_param(A.PARAM_LIST_ACTIVE))   # 1 = neighbour lists on the strips (ghost records by peer stores from k_step)
    if os.environ.get("STRIP_TEST_EXPECT_LISTS"):
        assert lists == int(os.environ["STRIP_TEST_EXPECT_LISTS"]), f"rank {rank}: list pipeline active={lists}"
    collisions = overflow = 0
    if os.environ.get("STRIP_TEST_IO") == "pipelined":
        # the frame loop of bench.py at N > 1 (pipelined distributed host I/O), with zero forces so that the world still has to
        # equal the undisturbed single world; the (slot, position) list that lands last must be this rank's owned bodies
        import bench

        tot = {"collisions": 0, "list_overflow": 0}

        def on_step(st):
            assert st["nan_detected"] == 0, st
            tot["collisions"] += st["collisions"]
            tot["list_overflow"] += st["list_overflow"]

        cap = sc.n_bodies
        pin = (lambda t: t) if os.environ.get("BLOBS_TEST_EMU") == "1" else (lambda t: t.pin_memory())
        sl = (pin(torch.zeros(cap, dtype=torch.int32)), pin(torch.zeros(cap, dtype=torch.int32)))
        xy = (pin(torch.zeros((cap, 2), dtype=torch.float32)), pin(torch.zeros((cap, 2), dtype=torch.float32)))
        cnt = (pin(torch.zeros(1, dtype=torch.int32)), pin(torch.zeros(1, dtype=torch.int32)))
        forces = pin(torch.zeros((cap, 2), dtype=torch.float32))
        n0 = w.read_owned_positions_ptr(sl[0].data_ptr(), xy[0].data_ptr(), cap)
        sl[1].copy_(sl[0])
        cnt[0][0] = cnt[1][0] = n0
        moved = bench.strip_pipelined_loop(w, steps, forces, sl, xy, cnt, cap, on_step=on_step)
        collisions, overflow = tot["collisions"], tot["list_overflow"]
        k = (steps - 1) & 1
        n_last = int(cnt[k][0])
        own_now = w.strip_owned().astype(bool)
        got_slots = sl[k].numpy()[:n_last].astype(np.int64)
        assert n_last == int(own_now.sum()) and len(np.unique(got_slots)) == n_last and own_now[got_slots].all(), "pipelined owned list != ownership"
        pos_now = w.read_positions()
        assert np.array_equal(xy[k].numpy()[:n_last].view(np.uint32), pos_now[got_slots].view(np.uint32)), "pipelined positions != device state"
        assert moved > 0
    else:
        for _ in range(steps):
            st = w.step(1 / 60)
            assert st["nan_detected"] == 0, st
            collisions += st["collisions"]
            overflow += st["list_overflow"]
    if os.environ.get("STRIP_TEST_EXPECT_LISTS_AFTER"):
        assert int(w.get_param(A.PARAM_LIST_ACTIVE)) == int(os.environ["STRIP_TEST_EXPECT_LISTS_AFTER"]), f"rank {rank}: list pipeline active after the run"
    own = w.strip_owned()
    bodies, _ = w.download_bodies()
    cols, _ = w.download_colliders()
    pack = np.concatenate([bodies["position"]["x"], bodies["position"]["y"], bodies["position_old"]["x"], bodies["position_old"]["y"],
                           bodies["calculated_velocity"]["x"], bodies["calculated_velocity"]["y"],
                           cols["desc"]["absolute_transform"]["translation"]["x"], cols["desc"]["absolute_transform"]["translation"]["y"]]).astype(np.float32)
    t_pack, t_own, t_own0 = torch.from_numpy(pack), torch.from_numpy(own.astype(np.uint8)), torch.from_numpy(own0.astype(np.uint8))
    t_col = torch.tensor([collisions], dtype=torch.int64)
    gp = [torch.zeros_like(t_pack) for _ in range(world)] if rank == 0 else None
    go = [torch.zeros_like(t_own) for _ in range(world)] if rank == 0 else None
    go0 = [torch.zeros_like(t_own0) for _ in range(world)] if rank == 0 else None
    dist.gather(t_pack, gp, 0)
    dist.gather(t_own, go, 0)
    dist.gather(t_own0, go0, 0)
    dist.all_reduce(t_col)
    ok = True
    if rank == 0:
        n = sc.n_bodies
        owners = np.stack([o.numpy() for o in go]).astype(np.int32)
        owners0 = np.stack([o.numpy() for o in go0]).astype(np.int32)
        assert (owners.sum(0) == 1).all() and (owners0.sum(0) == 1).all(), "every body must have exactly one owner"
        migrated = int((owners.argmax(0) != owners0.argmax(0)).sum())
        merged = np.zeros(8 * n, dtype=np.float32)
        for r in range(world):
            sel = np.tile(owners[r].astype(bool), 8)
            merged[sel] = gp[r].numpy()[sel]
        ref = blobs_b200.World(gravity=sc.gravity, device=local)
        S.build(ref, sc)
        ref_col = 0
        for _ in range(steps):
            ref_col += ref.step(1 / 60)["collisions"]
        rb, _ = ref.download_bodies()
        rc, _ = ref.download_colliders()
        want = np.concatenate([rb["position"]["x"], rb["position"]["y"], rb["position_old"]["x"], rb["position_old"]["y"],
                               rb["calculated_velocity"]["x"], rb["calculated_velocity"]["y"],
                               rc["desc"]["absolute_transform"]["translation"]["x"], rc["desc"]["absolute_transform"]["translation"]["y"]]).astype(np.float32)
        bad = np.nonzero(merged.view(np.uint32) != want.view(np.uint32))[0]
        print(f"[strip test] ranks={world} spheres={n} steps={steps} migrated={migrated} collisions strips={int(t_col)} single={ref_col} "
              f"list_overflow rank0={overflow} p2p={p2p} lists={lists} list_rebuilds={int(w.get_param(A.PARAM_LIST_REBUILDS))}/{int(w.get_param(A.PARAM_LIST_SUBSTEPS))} "
              f"mismatches={len(bad)}")
        ok = len(bad) == 0 and int(t_col) == ref_col and (world == 1 or migrated > 0)
        if len(bad):
            print("first mismatches (field*n + slot):", bad[:10], merged[bad[:10]], want[bad[:10]])
    if os.environ.get("STRIP_TEST_BENCH_PROBE") == "1":
        # bench.py's own N > 1 self-check (a SECOND strip world in this process, after the first one): must say bit-identical
        import bench

        w.close()
        knobs = {A.PARAM_STRIP_P2P: 1} if p2p else {}
        out = bench.strip_parity_probe(dist, rank, world, local, "cpu", steps=10, ny=48, cols_per_rank=24, params=knobs)
        if rank == 0:
            print("[strip test] bench probe:", out)
            ok = ok and out["bit_identical"] and out["migrated_bodies"] > 0 and out["list_pipeline_active"] == lists
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
