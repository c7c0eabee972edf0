#!/bin/bash
# Round-2, call 12 (4 GPUs): why do lists on strips lose at N = 8? graph replay on/off, grid for reference.
set -u
O=gpurun_out/r2l
mkdir -p $O
export PYTHONUNBUFFERED=1
run() {
  local name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus 4 --steps 20 --warmup 5 > $O/$name.json 2> $O/$name.err
  echo "$name rc=$?" >> $O/runs.log
}
run lists_graph BLOBS_B200_LIST=2
run lists_nograph BLOBS_B200_LIST=2 BLOBS_BENCH_GRAPH=0
run grid BLOBS_B200_LIST=0
ls -la $O
