"""Cycle breakdown of restart 0's tile 0 in the general-path kernels over a whole run (library
built with `python -m viloca_b200.build --profile`): where one restart's update spends its time."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from viloca_b200 import _capi as capi, engine, synthetic  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "hiv5k_uqs"
nw = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else None
idx = [int(x) for x in os.environ["VL_WINDOWS"].split(",")] if os.environ.get("VL_WINDOWS") else None   # e.g. VL_WINDOWS=80,83
wins = [w.pin() for w in synthetic.make_workload(name, nw, indices=idx)]
H = [("wait published", 6), ("prologue", 0), ("operand loop", 1), ("warp reduction", 2), ("sparse", 3),
     ("softmax epilogue", 4), ("sums", 5), ("signal", 7)]
Z = [("wait published", 14), ("wait haplotype tiles", 15), ("prologue", 8), ("operand loop", 9), ("sparse", 10),
     ("softmax epilogue", 11), ("records + ticket", 12), ("scalar update (last tile)", 13)]
with engine.CaviBatch(wins) as b:
    b.upload()
    buf = (C.c_uint64 * 36)()
    capi.check(capi.lib.vl_debug_counters(buf, 1))
    t0 = time.perf_counter(); b.run(); t1 = time.perf_counter()
    capi.check(capi.lib.vl_debug_counters(buf, 0))
    v = list(buf)[16:]
    nh, nz = max(v[16], 1), max(v[17], 1)
    print(f"run {1e3*(t1-t0):.1f} ms; restart 0: {v[16]} haplotype-tile / {v[17]} read-tile samples")
    for title, rows, n in (("haplotype tile 0", H, nh), ("read tile 0", Z, nz)):
        tot = sum(v[i] for _, i in rows)
        print(f" {title}: {tot/n:.0f} cycles = {tot/n/1965:.1f} us per update")
        for lab, i in rows:
            print(f"   {lab:28s} {v[i]/n:9.0f} cycles  {100*v[i]/max(tot,1):5.1f}%")
