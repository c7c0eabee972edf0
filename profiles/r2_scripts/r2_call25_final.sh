#!/bin/bash
# Round-2, final single-GPU call: the whole GPU suite, smoke, both bench arms as the driver runs them, fresh ncu evidence.
set -u
O=gpurun_out/r2x
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/runs.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/runs.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_driver.json 2> $O/bench_driver.err; echo "bench rc=$?" >> $O/runs.log
# plain launches for ncu: kernels that can set a conditional handle (k_step, k_nl_decide) are not profiled inside a replayed graph
BLOBS_BENCH_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file $O/ncu_launches_sparse.csv \
    python bench.py --steps 3 --warmup 10 --no-cpu-baseline --no-late > $O/ncu_launches_sparse.log 2>&1
BLOBS_BENCH_GRAPH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_step" -s 40 -c 1 -o $O/step_sparse -f \
    python bench.py --steps 2 --warmup 10 --no-cpu-baseline --no-late --no-flush > $O/ncu_step_sparse.log 2>&1
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?" >> $O/runs.log
timeout 300 python bench.py --workload cfg3 --warmup 5 --steps 20 --no-cpu-baseline > $O/cfg3_driver.json 2>> $O/err.log; echo "cfg3 rc=$?" >> $O/runs.log
timeout 300 python bench.py --workload cfg4 --warmup 30 --steps 30 --no-cpu-baseline > $O/cfg4.json 2>> $O/err.log; echo "cfg4 rc=$?" >> $O/runs.log
cat $O/runs.log; ls -la $O
