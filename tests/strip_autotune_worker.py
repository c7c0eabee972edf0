"""Stand-in for the parent ranks of bench.py at N > 1 (run under torchrun by test_bench_autotune.py): every rank calls
bench.autotune_strips, which starts one child per rank; the children must find each other on their own port, probe, and hand
every parent the same verdict."""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import bench  # noqa: E402

if __name__ == "__main__":
    args = argparse.Namespace(tune=0, workload="cfg2", warmup=200, steps=100)
    tune, p2p, rep = bench.autotune_strips(args)
    # one write() call including the newline, so that the two ranks' lines cannot interleave on the shared pipe
    sys.stdout.write("VERDICT " + json.dumps({"rank": int(os.environ["RANK"]), "tune": tune, "p2p": p2p, "report": rep}) + "\n")
    sys.stdout.flush()
