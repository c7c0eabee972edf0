#!/bin/bash
# Round-2, call 31 (1 GPU, the last seconds of the budget): joint / spring / API GPU tests with joint_advance on by default.
set -u
O=gpurun_out/r2ad
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 60 python -m pytest tests/test_gpu_parity.py tests/test_physics_api.py tests/test_rigid_body_api.py tests/test_hand_vectors.py -q -m gpu -k "soft or joint or spring or multi or api or hand or rigid" > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/runs.log
cat $O/runs.log; tail -3 $O/tests.log
