"""Per-window update counts of a synthetic workload (one run), written to
profiles/bench_iterations.json: the table the reference arm of bench.py and the longest-first
scheduler's iteration predictor use (trajectories are identical to the reference's: parity tests).

    python tools/iterations.py <workload> [n_windows]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from viloca_b200 import engine, synthetic  # noqa: E402

name = sys.argv[1]
nw = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else None
wins = [w.pin() for w in synthetic.make_workload(name, nw)]
with engine.CaviBatch(wins) as b:
    b.upload(); b.run()
    out = b.fetch(state="none", want_history=False)
    stats = b.problem_stats()
path = os.path.join(ROOT, "profiles", "bench_iterations.json")
table = json.load(open(path)) if os.path.exists(path) else {}
table[name] = {"n_updates": [r.n_updates for o in out for r in o.restarts], "n_reads": [w.n_reads for w in wins],
               "n_raw": [getattr(w, "n_raw", w.n_reads) for w in wins], "length": [w.length for w in wins],
               "final_clusters": [int(s[2]) for s in stats], "status": [r.status for o in out for r in o.restarts]}
json.dump(table, open(path, "w"))
u = table[name]["n_updates"]
print(name, "device ms", b.last_device_ms(), "updates: mean", sum(u) / len(u), "max", max(u))
print(sorted(range(len(u)), key=lambda i: -u[i])[:20], sorted(u)[-20:])
