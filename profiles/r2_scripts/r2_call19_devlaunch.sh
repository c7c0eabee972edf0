#!/bin/bash
# Round-2, call 19 (1 GPU): device-side launch of the list rebuild - probe, GPU parity suite, driver-window A/B.
set -u
O=gpurun_out/r2r
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 120 profiles/r2_scripts/cdp_probe > $O/cdp_probe.txt 2>&1; echo "probe rc=$?" >> $O/runs.log
cat $O/cdp_probe.txt
timeout 1200 python -m pytest tests -x -q -m gpu > $O/gpu_tests.log 2>&1; echo "tests rc=$?" >> $O/runs.log
tail -5 $O/gpu_tests.log
for v in 1 0; do
  BLOBS_B200_DEVLAUNCH=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-late --no-cpu-baseline > $O/driver_devlaunch$v.json 2> $O/driver_devlaunch$v.err; echo "bench devlaunch=$v rc=$?" >> $O/runs.log
  BLOBS_B200_DEVLAUNCH=$v timeout 600 python bench.py --steps 30 --warmup 60 --no-late --no-cpu-baseline > $O/sparse_devlaunch$v.json 2> $O/sparse_devlaunch$v.err; echo "sparse devlaunch=$v rc=$?" >> $O/runs.log
done
cat $O/runs.log
