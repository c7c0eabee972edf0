#!/bin/bash
# Round-2 strip-path measurements on TWO GPUs, in ONE gpurun call (charged 2x; ~6 box-minutes):
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash profiles/r2_second_call_n2.sh'
# Everything below was written after round 1's GPU budget was spent and has only run on the host-compiled build (tests/emu):
#  1. 2-GPU parity: NCCL exchange, peer-memory exchange (k_strip_push over CUDA IPC) + bench.py's pipelined frame loop
#  2. bench lines, 4 M spheres over 2 strips: exchange = NCCL / peer stores / peer stores + CUDA-graph replay, k_main vs k_tile
# Every rank runs under `timeout`: k_strip_push gives up after ~4 s on its own (bit 3 of nan_detected), nothing here can hang a GPU.
set -u
O=gpurun_out/r2n2
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
run() {  # name, extra env...
  local name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus 2 --steps 60 --warmup 60 ${TUNE:+--tune $TUNE} > $O/$name.json 2> $O/$name.err
  echo "$name rc=$?" >> $O/runs.log
}
run auto            BLOBS_BENCH_AUTOTUNE=1          # the default line: the child-group probe picks kernel variant and exchange kind
export BLOBS_BENCH_AUTOTUNE=0                        # everything below forces its combination
run nccl            BLOBS_B200_STRIP_P2P=0
run p2p             BLOBS_B200_STRIP_P2P=1
run p2p_graph       BLOBS_B200_STRIP_P2P=1 BLOBS_B200_STRIP_GRAPH=1
TUNE=11 run tile_nccl       BLOBS_B200_STRIP_P2P=0
TUNE=11 run tile_p2p        BLOBS_B200_STRIP_P2P=1
TUNE=11 run tile_p2p_graph  BLOBS_B200_STRIP_P2P=1 BLOBS_B200_STRIP_GRAPH=1
ls -la $O
