#!/bin/bash
# Round-2, call 17 (2 GPUs): multi-GPU parity tests with the tail publish, then lists (tail publish / separate publish kernel) and grid + graph replay.
set -u
O=gpurun_out/r2p
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > $O/tests_2gpu.log 2>&1; echo "tests rc=$?" >> $O/runs.log
run() {
  local name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus 2 --steps 20 --warmup 5 > $O/$name.json 2> $O/$name.err
  echo "$name rc=$?" >> $O/runs.log
}
run lists_tail BLOBS_X=1
run lists_pubkernel BLOBS_B200_NLS_TAIL=0
run grid_graph BLOBS_B200_LIST=0
tail -3 $O/tests_2gpu.log; cat $O/runs.log
