#!/bin/bash
# Round-2, call 20 (1 GPU): k_step occupancy variants on the driver window + a fresh full capture of k_step.
set -u
O=gpurun_out/r2s
mkdir -p $O
export PYTHONUNBUFFERED=1
for t in 0 1 2 3 4 5; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-late --no-cpu-baseline --no-flush --tune $t > $O/tune$t.json 2> $O/tune$t.err; echo "tune $t rc=$?" >> $O/runs.log
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_step" -s 40 -c 1 -o $O/kstep_final -f python bench.py --steps 2 --warmup 10 --no-cpu-baseline --no-flush --no-late > $O/ncu_f.log 2>&1
BLOBS_BENCH_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_step" -s 40 -c 1 -o $O/kstep_final_nograph -f python bench.py --steps 2 --warmup 10 --no-cpu-baseline --no-flush --no-late > $O/ncu_f2.log 2>&1
cat $O/runs.log; ls -la $O
