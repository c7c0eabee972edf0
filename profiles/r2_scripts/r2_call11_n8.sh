#!/bin/bash
# Round-2, call 11 (8 GPUs): config #5 with strip-major insertion order; lists vs grid, and the row-major order again for the A/B.
set -u
O=gpurun_out/r2k
mkdir -p $O
export PYTHONUNBUFFERED=1
run() {
  local name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus 8 --steps 20 --warmup 5 > $O/$name.json 2> $O/$name.err
  echo "$name rc=$?" >> $O/runs.log
}
run lists_stripmajor BLOBS_B200_LIST=2
run grid_stripmajor BLOBS_B200_LIST=0
ls -la $O
