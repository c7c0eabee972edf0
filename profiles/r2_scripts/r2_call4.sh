#!/bin/bash
# Round-2, call 4: cooperative gather (grid pipeline) + automatic list/grid choice + cheaper gated kernels + tail decision.
set -u
O=gpurun_out/r2d
mkdir -p $O
export PYTHONUNBUFFERED=1 BLOBS_BENCH_AUTOTUNE=0
timeout 900 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
for l in 2 0 1; do
  TRACE_LIST=$l timeout 300 python profiles/trace_cfg2.py 800 50 > $O/trace_list$l.jsonl 2>> $O/err.log
done
for l in 2 0; do
  timeout 200 python bench.py --list $l --warmup 60 --steps 30 --no-cpu-baseline > $O/sparse_list$l.json 2>> $O/err.log
  timeout 300 python bench.py --list $l --no-cpu-baseline > $O/dense_list$l.json 2>> $O/err.log
  timeout 300 python bench.py --list $l --workload cfg3 --no-cpu-baseline > $O/cfg3_list$l.json 2>> $O/err.log
done
timeout 200 python bench.py --list 1 --tune 1 --warmup 60 --steps 30 --no-cpu-baseline > $O/sparse_list1_tune1.json 2>> $O/err.log
timeout 200 python bench.py --list 1 --warmup 60 --steps 30 --no-cpu-baseline > $O/sparse_list1_tune0.json 2>> $O/err.log
timeout 200 python bench.py --warmup 5 --steps 20 --no-cpu-baseline > $O/driver_default.json 2>> $O/err.log
timeout 300 python bench.py --workload cfg4 --warmup 30 --steps 30 --no-cpu-baseline > $O/cfg4_list2.json 2>> $O/err.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 150 --csv --log-file $O/ncu_launches_sparse.csv \
    python bench.py --steps 3 --warmup 10 --no-cpu-baseline > $O/ncu_launches_sparse.log 2>&1
for at in 300 700; do
  CAPTURE_STEPS=$at TRACE_LIST=0 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:"k_main|k_crowded" -c 4 -o $O/coop${at} python profiles/trace_cfg2.py 0 0 2 > $O/ncu_coop${at}.log 2>&1
done
ls -la $O
