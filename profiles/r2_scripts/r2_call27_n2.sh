#!/bin/bash
# Round-2, call 27 (2 GPUs): the strip bench line with the final bench.py (e2e: 2 untimed steps + median of 3 windows).
set -u
O=gpurun_out/r2z
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 20 --warmup 5 > $O/lists.json 2> $O/lists.err; echo "lists rc=$?" >> $O/runs.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/reference.json 2> $O/reference.err; echo "reference rc=$?" >> $O/runs.log
cat $O/runs.log
