"""Small mixed batch (uqs + lep, three restarts, a split-tile window) for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_target.py
    compute-sanitizer --tool racecheck python tools/sanitize_target.py
"""
import dataclasses
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from viloca_b200 import engine, synthetic  # noqa: E402

base = synthetic.WORKLOADS["hiv5k_uqs"]
wins = [synthetic.make_window(dataclasses.replace(base, n_clusters=24, called_length=60, n_starts=3), 1, n_reads=150, length=80, max_iter=40),
        synthetic.make_window(dataclasses.replace(base, mode="lep", n_clusters=16, called_length=50), 2, n_reads=120, length=64, max_iter=40),
        synthetic.make_window(dataclasses.replace(base, n_clusters=12, called_length=None), 3, n_reads=1400, length=64, max_iter=24)]
out = engine.run_windows(wins)
print([(r.n_updates, r.status) for o in out for r in o.restarts])
