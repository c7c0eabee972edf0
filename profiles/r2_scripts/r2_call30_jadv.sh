#!/bin/bash
# Round-2, call 30 (1 GPU, last GPU-minutes): k_joints_fused advancing its bodies itself (BLOBS_B200_JADV=1) vs the separate k_integrate pass, config #4.
set -u
O=gpurun_out/r2ac
mkdir -p $O
export PYTHONUNBUFFERED=1
BLOBS_B200_JADV=1 timeout 100 python bench.py --workload cfg4 --warmup 30 --steps 30 --no-cpu-baseline > $O/cfg4_jadv1.json 2> $O/err1.log; echo "jadv1 rc=$?" >> $O/runs.log
BLOBS_B200_JADV=1 timeout 100 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "soft or joint" > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/runs.log
BLOBS_B200_JADV=0 timeout 100 python bench.py --workload cfg4 --warmup 30 --steps 30 --no-cpu-baseline > $O/cfg4_jadv0.json 2> $O/err0.log; echo "jadv0 rc=$?" >> $O/runs.log
cat $O/runs.log; tail -2 $O/tests.log
