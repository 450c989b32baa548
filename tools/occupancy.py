"""Residency of the fused chunk kernel per kernel class, from the CUDA occupancy calculator."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from viloca_b200 import _capi as capi  # noqa: E402

for mode, label in ((0, "uqs"), (1, "lep")):
    for cls, width in enumerate((2, 4, 8, 16)):
        for deep in (0, 1):
            c, s, r = C.c_int32(0), C.c_int32(0), C.c_int32(0)
            capi.check(capi.lib.vl_debug_occupancy(0, cls, mode, deep, C.byref(c), C.byref(s), C.byref(r)))
            print(f"{label} class {cls} (<= {width:2d} cluster tiles) {'deep' if deep else 'std '}: "
                  f"{c.value} CTAs/SM, {s.value} B shared per CTA, {r.value} registers")
