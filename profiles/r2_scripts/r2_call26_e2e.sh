#!/bin/bash
# Round-2, call 26 (1 GPU): the driver's bench line with the e2e region on a fresh world, median of 3 windows.
set -u
O=gpurun_out/r2y
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_driver.json 2> $O/bench_driver.err; echo "bench rc=$?" >> $O/runs.log
timeout 300 python bench.py --workload cfg3 --warmup 5 --steps 20 --no-cpu-baseline > $O/cfg3_driver.json 2>> $O/err.log; echo "cfg3 rc=$?" >> $O/runs.log
cat $O/runs.log
