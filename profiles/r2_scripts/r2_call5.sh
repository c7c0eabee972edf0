#!/bin/bash
# Round-2, call 5: cooperative gather v2, new bench.py, full-size parity tests.
set -u
O=gpurun_out/r2e
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
timeout 600 python bench.py --warmup 5 --steps 20 > $O/bench_driver.json 2> $O/bench_driver.err
timeout 600 python bench.py --impl reference --warmup 5 --steps 20 > $O/bench_reference.json 2> $O/bench_reference.err
TRACE_LIST=2 timeout 300 python profiles/trace_cfg2.py 800 50 > $O/trace_list2.jsonl 2>> $O/err.log
timeout 300 python bench.py --workload cfg3 --no-cpu-baseline > $O/cfg3.json 2>> $O/err.log
timeout 300 python bench.py --workload cfg3 --warmup 5 --steps 20 --no-cpu-baseline > $O/cfg3_early.json 2>> $O/err.log
for at in 300 700; do
  CAPTURE_STEPS=$at TRACE_LIST=0 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:"k_main" -c 2 -o $O/coop${at} python profiles/trace_cfg2.py 0 0 2 > $O/ncu_coop${at}.log 2>&1
done
ls -la $O
