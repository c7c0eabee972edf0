#!/bin/bash
# Round-2, call 8: conditional (IF) graph nodes around the rebuild kernels: parity + A/B against self-gating launches.
set -u
O=gpurun_out/r2h
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cfg1 or full_size or every_broadphase or cuda_graph or batched or soft_blobs" > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
for c in 1 0; do
  BLOBS_B200_COND=$c timeout 300 python bench.py --warmup 5 --steps 20 --no-cpu-baseline --no-late > $O/driver_cond$c.json 2>> $O/err.log
  BLOBS_B200_COND=$c timeout 300 python bench.py --warmup 60 --steps 30 --no-cpu-baseline --no-late > $O/sparse_cond$c.json 2>> $O/err.log
done
timeout 300 python bench.py --workload cfg3 --warmup 5 --steps 20 --no-cpu-baseline > $O/cfg3_early.json 2>> $O/err.log
timeout 300 python bench.py --workload cfg4 --warmup 30 --steps 30 --no-cpu-baseline > $O/cfg4.json 2>> $O/err.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file $O/ncu_launches_sparse.csv \
    python bench.py --steps 3 --warmup 10 --no-cpu-baseline --no-late > $O/ncu_launches_sparse.log 2>&1
ls -la $O
