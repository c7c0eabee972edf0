#!/bin/bash
# Round-2, call 7: full GPU suite + skin sweep + tune A/B on one GPU.
set -u
O=gpurun_out/r2g
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
for sk in 0.4 0.6 0.8; do
  timeout 300 python bench.py --warmup 5 --steps 20 --skin $sk --no-cpu-baseline --no-late > $O/driver_skin$sk.json 2>> $O/err.log
  timeout 300 python bench.py --warmup 60 --steps 30 --skin $sk --no-cpu-baseline --no-late > $O/sparse_skin$sk.json 2>> $O/err.log
done
timeout 300 python bench.py --warmup 200 --steps 50 --list 0 --tune 2 --no-cpu-baseline --no-late > $O/dense_batch2.json 2>> $O/err.log
timeout 300 python bench.py --warmup 200 --steps 50 --list 0 --no-cpu-baseline --no-late > $O/dense_batch4.json 2>> $O/err.log
timeout 300 python bench.py --warmup 700 --steps 50 --list 0 --tune 2 --no-cpu-baseline --no-late > $O/late_batch2.json 2>> $O/err.log
timeout 300 python bench.py --warmup 700 --steps 50 --list 0 --no-cpu-baseline --no-late > $O/late_batch4.json 2>> $O/err.log
ls -la $O
