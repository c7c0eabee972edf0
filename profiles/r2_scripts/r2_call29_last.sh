#!/bin/bash
# Round-2, last call (1 GPU): the whole GPU suite and smoke on the code as committed.
set -u
O=gpurun_out/r2ab
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/runs.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/runs.log
tail -3 $O/tests.log; cat $O/smoke.log $O/runs.log
