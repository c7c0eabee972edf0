#!/bin/bash
# Round-2, call 3: --set full captures of the GRID pipeline's kernels (k_main<POOLED>, k_crowded, k_scan, k_scatter) and of the list
# pipeline's (k_step, k_nl_*) in the compressing (step 300) and settled (step 700) states of config #2.
set -u
O=gpurun_out/r2c
mkdir -p $O
export PYTHONUNBUFFERED=1
for at in 300 700; do
  for l in 0 1; do
    CAPTURE_STEPS=$at TRACE_LIST=$l timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:"k_main|k_crowded|k_scan|k_scatter|k_step|k_nl_" -c 7 -o $O/dense${at}_list$l python profiles/trace_cfg2.py 0 0 2 > $O/ncu_${at}_list$l.log 2>&1
  done
done
ls -la $O
