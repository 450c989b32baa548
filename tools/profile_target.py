"""Short, deterministic profiling target: upload the benchmark workload and run a few CAVI
iterations (all restarts active).  Used under ncu; prints nothing that is a bench value."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from viloca_b200 import engine, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="hiv5k_uqs")
ap.add_argument("--windows", type=int, default=145)
ap.add_argument("--updates", type=int, default=6)
ap.add_argument("--reads", type=int, default=None, help="raw reads per window (default: the workload's coverage)")
ap.add_argument("--run", action="store_true", help="vl_batch_run (fused chunk kernel) capped at --updates iterations")
args = ap.parse_args()
spec = synthetic.WORKLOADS[args.workload]
wins = [synthetic.make_window(spec, i, n_reads=args.reads, max_iter=args.updates if args.run else 50000)
        for i in range(args.windows)]
with engine.CaviBatch(wins) as b:
    b.upload()
    if args.run:
        b.run()
    else:
        b.step(args.updates)
    out = b.fetch(state="none")
print("ok", len(wins), "windows", sum(o.restarts[0].n_updates for o in out), "updates")
