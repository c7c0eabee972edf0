#!/bin/bash
# Round-2, call 23 (8 GPUs): config #5 (16 777 216 spheres) with the final code: library defaults (lists on strips), then the grid pipeline.
set -u
O=gpurun_out/r2v
mkdir -p $O
export PYTHONUNBUFFERED=1
run() {
  local name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus 8 --steps 20 --warmup 5 > $O/$name.json 2> $O/$name.err
  echo "$name rc=$?" >> $O/runs.log
}
run default BLOBS_X=1
run grid BLOBS_B200_LIST=0
cat $O/runs.log
