#!/bin/bash
# Round-2, call 16 (1 GPU): strip-rank diagnostics at 2 M spheres per GPU.
set -u
O=gpurun_out/r2o
mkdir -p $O
export PYTHONUNBUFFERED=1
for m in plain strip; do for l in 1 0; do
  timeout 300 python profiles/r2_scripts/strip_diag.py $m $l >> $O/diag.jsonl 2>> $O/diag.err
done; done
cat $O/diag.jsonl
# launch lists (cold-cache, serialised): plain and strip, list pipeline
for m in plain strip; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 160 --csv --log-file $O/launches_${m}_lists.csv python profiles/r2_scripts/strip_diag.py $m 1 2 > $O/ncu_l_$m.log 2>&1
done
# full captures of k_step: plain vs strip
for m in plain strip; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_step" -s 40 -c 2 -o $O/kstep_$m -f python profiles/r2_scripts/strip_diag.py $m 1 2 > $O/ncu_f_$m.log 2>&1
done
ls -la $O
