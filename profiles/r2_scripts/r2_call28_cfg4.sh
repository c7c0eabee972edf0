#!/bin/bash
# Round-2, call 28 (1 GPU): joint kernels without the two atan2f per joint (supplementary-angle identity): parity tests with joints, cfg4.
set -u
O=gpurun_out/r2aa
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_physics_api.py tests/test_rigid_body_api.py -q -m gpu -k "joint or soft or blob or matrix or spring or cfg4 or api" > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/runs.log
tail -3 $O/tests.log
timeout 300 python bench.py --workload cfg4 --warmup 30 --steps 30 --no-cpu-baseline > $O/cfg4.json 2>> $O/err.log; echo "cfg4 rc=$?" >> $O/runs.log
cat $O/runs.log
