#!/bin/bash
# Round-2, call 6 (2 GPUs): strip-decomposed world on the list pipeline vs the grid pipeline.
set -u
O=gpurun_out/r2f
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
run() {  # name, extra args / env...
  local name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus 2 --steps ${STEPS:-20} --warmup ${WARM:-5} > $O/$name.json 2> $O/$name.err
  echo "$name rc=$?" >> $O/runs.log
}
run lists_driver    BLOBS_B200_LIST=2
run grid_p2p_driver BLOBS_B200_LIST=0
run grid_nccl_driver BLOBS_B200_LIST=0 BLOBS_B200_STRIP_P2P=0
STEPS=60 WARM=60 run lists_60 BLOBS_B200_LIST=2
STEPS=60 WARM=60 run grid_p2p_60 BLOBS_B200_LIST=0
STEPS=100 WARM=200 run lists_auto_200 BLOBS_B200_LIST=2
ls -la $O
