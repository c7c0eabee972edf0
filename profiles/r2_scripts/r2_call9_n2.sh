#!/bin/bash
# Round-2, call 9 (2 GPUs): strips on lists after the rebuild-enumeration change, skin 0.6.
set -u
O=gpurun_out/r2i
mkdir -p $O
export PYTHONUNBUFFERED=1
run() {
  local name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus 2 --steps ${STEPS:-20} --warmup ${WARM:-5} > $O/$name.json 2> $O/$name.err
  echo "$name rc=$?" >> $O/runs.log
}
run lists_driver    BLOBS_B200_LIST=2
run lists_driver_skin08 BLOBS_B200_LIST=2 BLOBS_B200_SKIN=0.8
run grid_p2p_driver BLOBS_B200_LIST=0
STEPS=60 WARM=60 run auto_60 BLOBS_B200_LIST=2
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
ls -la $O
