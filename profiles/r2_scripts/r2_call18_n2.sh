#!/bin/bash
# Round-2, call 18 (2 GPUs): lists (tail publish / separate publish kernel) and grid + graph replay; bench line now carries parity_vs_single_gpu.
set -u
O=gpurun_out/r2q
mkdir -p $O
export PYTHONUNBUFFERED=1
run() {
  local name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus 2 --steps 20 --warmup 5 > $O/$name.json 2> $O/$name.err
  echo "$name rc=$?" >> $O/runs.log
}
run lists_tail BLOBS_X=1
run lists_pubkernel BLOBS_B200_NLS_TAIL=0
run grid_graph BLOBS_B200_LIST=0
cat $O/runs.log
