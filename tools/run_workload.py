"""One upload + run + fetch of a synthetic workload (profiling target for ncu).

    python tools/run_workload.py [workload] [n_windows] [repeats]

Environment: VL_N_READS / VL_MAX_ITER override the raw reads per window and the iteration cap
(bounded samples of the large workloads), VL_FUSED=0 selects the per-phase launches.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from viloca_b200 import engine, synthetic  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "hiv5k_uqs"
nw = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else None
rep = int(sys.argv[3]) if len(sys.argv) > 3 else 1
kw = {}
if os.environ.get("VL_N_READS"):
    kw["n_reads"] = int(os.environ["VL_N_READS"])
if os.environ.get("VL_MAX_ITER"):
    kw["max_iter"] = int(os.environ["VL_MAX_ITER"])
idx = [int(x) for x in os.environ["VL_WINDOWS"].split(",")] if os.environ.get("VL_WINDOWS") else None   # e.g. VL_WINDOWS=80,83
wins = [w.pin() for w in synthetic.make_workload(name, nw, indices=idx, procs=min(32, os.cpu_count() or 1), **kw)]
print(f"{name}: {len(wins)} windows, N_unique {[w.n_reads for w in wins[:4]]}, L {wins[0].length}, restarts {wins[0].n_starts}",
      flush=True)
fused = os.environ.get("VL_FUSED", "1") != "0"
with engine.CaviBatch(wins, fused_chunks=fused) as b:
    for _ in range(rep):
        t0 = time.perf_counter(); b.upload(); t1 = time.perf_counter()
        n = b.run(); t2 = time.perf_counter()
        out = b.fetch(pinned=True); t3 = time.perf_counter()
        print(f"upload {1e3*(t1-t0):.1f} ms  run {1e3*(t2-t1):.1f} ms (device {b.last_device_ms():.1f} ms, {n} launches)"
              f"  fetch {1e3*(t3-t2):.1f} ms  windows {len(wins)}", flush=True)
