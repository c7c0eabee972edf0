"""Where does a strip rank's time go? The per-GPU share of the N > 1 bench worlds (512 x 4096 = 2 097 152 spheres) on ONE GPU, as a
plain world and as a one-rank strip world (same kernels as a real strip rank - owned-list enumeration, STRIP templates, publish /
decide - but no peer), with either pipeline. Usage: python profiles/r2_scripts/strip_diag.py <plain|strip> <list 0|1> [steps]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import blobs_b200  # noqa: E402
from blobs_b200 import scenes as S  # noqa: E402

A = blobs_b200.abi
mode, lst = sys.argv[1], int(sys.argv[2])
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
DT = 1 / 60
nx, ny = 512, 4096
sc = S.lattice_scene(nx, ny, 1.05, (0.0, 0.0), 1, 0.5, 0.5, jitter=0.04, vel_disc=1.0, constraint_r=0.8 * ny, name="one-strip", cell_size=1.0)
w = blobs_b200.World(gravity=sc.gravity, device=0, body_capacity=sc.n_bodies, collider_capacity=sc.n_colliders)
S.build(w, sc)
w.set_param(A.PARAM_LIST, lst)
if os.environ.get("DIAG_TUNE"):
    w.set_param(A.PARAM_TUNE, int(os.environ["DIAG_TUNE"]))
if mode == "strip":
    w.set_param(A.PARAM_STRIP_P2P, 1)
    w.strip_configure(0, 1, float("-inf"), float("inf"), blobs_b200.World.strip_unique_id(), ghost_capacity=4 * ny, migrate_capacity=2 * ny)
n = sc.n_bodies
w.step(DT, n=5)
st = w.step(DT, n=K)
ms = st["gpu_ms"] / K
w.profile_enable(True)
for _ in range(K):
    w.step(DT)
prof = w.profile_read()
w.profile_enable(False)
w.set_param(A.PARAM_GRAPH, 0)
w.profile_enable(2)
for _ in range(K):
    w.step(DT)
pm = w.profile_read()
w.profile_enable(False)
print(json.dumps({"tune": int(os.environ.get("DIAG_TUNE", "0")), "mode": mode, "list": lst, "list_active": int(w.get_param(A.PARAM_LIST_ACTIVE)), "spheres": n, "ms_per_step": ms, "value": n / (ms / 1e3),
                  "main_avg_launch_us": 1e3 * pm["main"][0] / max(pm["main"][1], 1),
                  "kernel_ms_per_step": {k: round(v[0] / K, 4) for k, v in prof.items() if v[1]},
                  "rebuilds": [w.get_param(A.PARAM_LIST_REBUILDS), w.get_param(A.PARAM_LIST_SUBSTEPS)]}), flush=True)
w.close()
