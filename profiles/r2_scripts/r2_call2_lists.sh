#!/bin/bash
# Round-2, call 2: first run of the neighbour-list pipeline (k_step + gated rebuild kernels) on a B200.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/r2_call2_lists.sh'
set -u
O=gpurun_out/r2b
mkdir -p $O
export PYTHONUNBUFFERED=1 BLOBS_BENCH_AUTOTUNE=0
timeout 900 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
for l in 1 0; do
  TRACE_LIST=$l timeout 300 python profiles/trace_cfg2.py 800 50 > $O/trace_list$l.jsonl 2>> $O/err.log
  timeout 200 python bench.py --list $l --warmup 60 --steps 30 --no-cpu-baseline > $O/sparse_list$l.json 2>> $O/err.log
  timeout 300 python bench.py --list $l --no-cpu-baseline > $O/dense_list$l.json 2>> $O/err.log
  timeout 300 python bench.py --list $l --workload cfg3 --no-cpu-baseline > $O/cfg3_list$l.json 2>> $O/err.log
  timeout 300 python bench.py --list $l --workload cfg4 --warmup 30 --steps 30 --no-cpu-baseline > $O/cfg4_list$l.json 2>> $O/err.log
done
for sk in 0.2 0.8; do
  TRACE_LIST=1 TRACE_SKIN=$sk timeout 300 python profiles/trace_cfg2.py 800 100 > $O/trace_list1_skin$sk.jsonl 2>> $O/err.log
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file $O/ncu_launches_sparse.csv \
    python bench.py --steps 3 --warmup 10 --no-cpu-baseline > $O/ncu_launches_sparse.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_step|k_nl_build" -s 80 -c 2 -o $O/step_sparse \
    python bench.py --steps 2 --warmup 10 --no-cpu-baseline --no-flush > $O/ncu_step_sparse.log 2>&1
CAPTURE_STEPS=302 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_step|k_nl_build|k_crowded" --launch-skip 4500 --launch-count 4 \
    -o $O/step_dense python profiles/trace_cfg2.py 0 0 2 > $O/ncu_dense.log 2>&1
ls -la $O
