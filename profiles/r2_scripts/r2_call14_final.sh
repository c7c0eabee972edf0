#!/bin/bash
# Round-2, final single-GPU call: the whole GPU suite, smoke, both bench arms as the driver runs them, fresh ncu evidence.
set -u
O=gpurun_out/r2m
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_driver.json 2> $O/bench_driver.err; echo "bench rc=$?" >> $O/runs.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?" >> $O/runs.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?" >> $O/runs.log
timeout 300 python bench.py --workload cfg3 --warmup 5 --steps 20 --no-cpu-baseline > $O/cfg3_driver.json 2>> $O/err.log
timeout 300 python bench.py --workload cfg4 --warmup 30 --steps 30 --no-cpu-baseline > $O/cfg4.json 2>> $O/err.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file $O/ncu_launches_sparse.csv \
    python bench.py --steps 3 --warmup 10 --no-cpu-baseline --no-late > $O/ncu_launches_sparse.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_step" -s 40 -c 2 -o $O/step_sparse \
    python bench.py --steps 2 --warmup 10 --no-cpu-baseline --no-late --no-flush > $O/ncu_step_sparse.log 2>&1
CAPTURE_STEPS=700 TRACE_LIST=2 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"k_main" -c 1 -o $O/coop700 python profiles/trace_cfg2.py 0 0 2 > $O/ncu_coop700.log 2>&1
ls -la $O
