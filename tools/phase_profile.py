"""Per-kernel device time of one CAVI update at several points of a run (CUDA events around
every launch, vl_batch_profile_step), with the number of restarts still running and the mean
layout width: shows where the time of a whole run goes as clusters die and windows converge.

    python tools/phase_profile.py [workload] [n_windows]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from viloca_b200 import engine, synthetic  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "hiv5k_uqs"
    nw = int(sys.argv[2]) if len(sys.argv) > 2 else None
    wins = synthetic.make_workload(name, nw)
    rows = []
    with engine.CaviBatch(wins) as b:
        b.upload()
        done = 0
        for upto in (0, 8, 32, 128, 512, 1024, 2048, 3072, 4096, 6144):
            if upto > done:
                t0 = time.perf_counter()
                b.step(upto - done)
                dt = time.perf_counter() - t0
                done = upto
            else:
                dt = 0.0
            st = b.problem_stats()
            run = st[:, 0] == 1
            if not run.any():
                break
            prof = b.profile_step(4)
            done += 4
            row = {"update": upto, "running": int(run.sum()), "mean_slots": float(st[run, 2].mean()),
                   "mean_tiles": float(st[run, 1].mean()),
                   "haplo_us": 1e3 * prof["haplo"][0] / 4, "cluster_us": 1e3 * prof["cluster"][0] / 4,
                   "launches_per_update": sum(v[1] for v in prof.values()) / 4,
                   "step_wall_us_per_update": 1e6 * dt / max(upto, 1) if dt else None}
            rows.append(row)
            print(json.dumps(row), flush=True)
    return rows


if __name__ == "__main__":
    main()
