#!/bin/bash
# Round-2, call 24 (1 GPU): number of speculatively fetched list rows (NL_SPEC 2 / 3 / 4 / 6) x 5 or 6 CTAs per SM, driver window.
set -u
O=gpurun_out/r2w
mkdir -p $O
export PYTHONUNBUFFERED=1
for sp in 4 2 3 6; do for t in 2 3; do
  lib=""; [ $sp != 4 ] && lib=$PWD/build_variants/libvar_spec$sp.so
  BLOBS_B200_LIBRARY=$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-late --no-cpu-baseline --no-flush --tune $t > $O/spec${sp}_tune$t.json 2> $O/spec${sp}_tune$t.err; echo "spec $sp tune $t rc=$?" >> $O/runs.log
done; done
cat $O/runs.log
