"""Cycle breakdown of one resident CTA over a whole run (library built with --profile)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from viloca_b200 import _capi as capi, engine, synthetic  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "hiv5k_uqs"
nw = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else None
wins = [w.pin() for w in synthetic.make_workload(name, nw)]
labels = ["T sweep", "T->smem", "haplo epilogue", "pass1 dmma", "pass1 sparse", "pass1 softmax", "bar a", "pass2 dmma",
          "bar b", "reduce", "scalar", "", "", "", "iterations", "sum kt"]
with engine.CaviBatch(wins, resident=True) as b:
    b.upload()
    buf = (C.c_uint64 * 36)()
    capi.check(capi.lib.vl_debug_counters(buf, 1))
    t0 = time.perf_counter(); b.run(); t1 = time.perf_counter()
    capi.check(capi.lib.vl_debug_counters(buf, 0))
    vals = list(buf)
    its = max(vals[14], 1)
    print(f"run {1e3*(t1-t0):.1f} ms; CTA 0 of every resident launch: {vals[14]} iterations, mean kt {vals[15]/its:.2f}")
    tot = sum(vals[:11])
    for lab, v in zip(labels[:11], vals[:11]):
        print(f"  {lab:16s} {v/its:10.0f} cycles/iter  {100*v/tot:5.1f}%")
    print(f"  total            {tot/its:10.0f} cycles/iter = {tot/its/1.965e3:.1f} us at 1965 MHz")
