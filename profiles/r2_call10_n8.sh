#!/bin/bash
# Round-2, call 10 (8 GPUs): config #5, 16 777 216 spheres, driver window.
set -u
O=gpurun_out/r2j
mkdir -p $O
export PYTHONUNBUFFERED=1
run() {
  local name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus 8 --steps 20 --warmup 5 > $O/$name.json 2> $O/$name.err
  echo "$name rc=$?" >> $O/runs.log
}
run lists_driver BLOBS_B200_LIST=2
run grid_driver BLOBS_B200_LIST=0
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
ls -la $O
