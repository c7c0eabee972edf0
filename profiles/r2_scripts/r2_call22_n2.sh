#!/bin/bash
# Round-2, call 22 (2 GPUs): multi-GPU parity tests and the strip bench with the final k_step (no mid-kernel call, 6 CTAs per SM on strips).
set -u
O=gpurun_out/r2u
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > $O/tests_2gpu.log 2>&1; echo "tests rc=$?" >> $O/runs.log
run() {
  local name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus 2 --steps 20 --warmup 5 > $O/$name.json 2> $O/$name.err
  echo "$name rc=$?" >> $O/runs.log
}
run lists BLOBS_X=1
run lists_tune2 BLOBS_B200_TUNE=2
tail -3 $O/tests_2gpu.log; cat $O/runs.log
