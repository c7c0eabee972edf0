#!/bin/bash
# Round-2, call 21 (1 GPU): k_step without the mid-kernel overflow call: GPU parity suite, occupancy variants (plain 1 M driver window;
# one-rank strip world at 2 M).
set -u
O=gpurun_out/r2t
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -x -q -m gpu > $O/gpu_tests.log 2>&1; echo "tests rc=$?" >> $O/runs.log
tail -3 $O/gpu_tests.log
for t in 0 2 3 4 5; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-late --no-cpu-baseline --no-flush --tune $t > $O/tune$t.json 2> $O/tune$t.err; echo "tune $t rc=$?" >> $O/runs.log
done
for t in 0 3 5; do
  DIAG_TUNE=$t timeout 300 python profiles/r2_scripts/strip_diag.py strip 1 >> $O/diag.jsonl 2>> $O/diag.err
done
cat $O/diag.jsonl $O/runs.log
