"""Throughput of BASELINE config #2 as a function of SIMULATED time (one line per block of steps): the block falls
freely (few contacts), reaches the circle constraint at ~1.9 s, gets compressed against it (dense contacts, bodies with
more than 24 contributions) and settles. Usage: python profiles/trace_cfg2.py [total_steps] [block] [crowded_mode]
Also used as the workload for the dense-regime ncu captures (env CAPTURE_STEPS=N: run N steps, then one more step inside a
cudaProfilerStart/Stop window, for `ncu --profile-from-start off`). TRACE_LIST / TRACE_SKIN select the broadphase pipeline."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blobs_b200  # noqa: E402
from blobs_b200 import _abi as A, scenes  # noqa: E402

DT = 1.0 / 60.0


def main():
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    block = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    crowded = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    sc = scenes.cfg2(seed=1)
    w = blobs_b200.World(gravity=sc.gravity)
    scenes.build(w, sc)
    w.set_param(A.PARAM_CROWDED, crowded)
    if "TRACE_LIST" in os.environ:
        w.set_param(A.PARAM_LIST, int(os.environ["TRACE_LIST"]))
    if "TRACE_SKIN" in os.environ:
        w.set_param(A.PARAM_SKIN, float(os.environ["TRACE_SKIN"]))
    cap = int(os.environ.get("CAPTURE_STEPS", "0"))
    if cap:
        # ncu --profile-from-start off: run `cap` steps at full speed (CUDA graphs), then open the profiler window around one
        # more step launched kernel by kernel
        import torch

        w.step(DT, n=cap)
        w.set_param(A.PARAM_GRAPH, 0)
        w.step(DT)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        w.step(DT)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    n = sc.n_bodies
    done = 0
    while done < total:
        ms = 0.0
        coll = over = 0
        r0, s0 = w.get_param(A.PARAM_LIST_REBUILDS), w.get_param(A.PARAM_LIST_SUBSTEPS)
        for _ in range(block):
            st = w.step(DT)
            ms += st["gpu_ms"]
            coll += st["collisions"]
            over += st["list_overflow"]
        done += block
        print(json.dumps({"steps": [done - block, done], "sim_time_s": round(done * DT, 3), "ms_per_step": ms / block,
                          "sphere_steps_per_s": n * block / (ms / 1e3), "contacts_per_step": coll / block,
                          "list_overflow_per_step": over / block,
                          "rebuilds_per_substep": (w.get_param(A.PARAM_LIST_REBUILDS) - r0) / max(w.get_param(A.PARAM_LIST_SUBSTEPS) - s0, 1.0), "grid": [w.kernel_info()["grid_w"], w.kernel_info()["grid_h"]]}), flush=True)


if __name__ == "__main__":
    main()
