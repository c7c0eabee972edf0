#!/bin/bash
# Round-2, call 15 (8 GPUs): grid pipeline on strips with CUDA-graph replay (peer-store exchange), and config #3 on 8 GPUs.
set -u
O=gpurun_out/r2n
mkdir -p $O
export PYTHONUNBUFFERED=1
run() {
  local name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus 8 --steps 20 --warmup 5 ${EXTRA:-} > $O/$name.json 2> $O/$name.err
  echo "$name rc=$?" >> $O/runs.log
}
run grid_graph BLOBS_B200_LIST=0 BLOBS_B200_STRIP_GRAPH=1
EXTRA="--workload cfg3" run cfg3_n8
ls -la $O
