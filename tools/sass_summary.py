"""Per-kernel SASS evidence of the built library (runs without a GPU): registers, spills, shared
memory (cuobjdump -res-usage) and the counts of the instructions that matter here -- DMMA (FP64
tensor), UBLKCP (TMA bulk copy), SYNCS (mbarrier), and the tcgen05 family (UTCMMA / LDTM: none
expected, tcgen05 has no f64 kind).

    python tools/sass_summary.py [library] > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "viloca_b200", "lib", "libviloca_b200.so")
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
name = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        name = m.group(1)
        continue
    if name and "REG:" in line:
        usage[name] = dict(re.findall(r"(\w+):(\d+)", line))
        name = None
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.defaultdict(collections.Counter)
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur:
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for key in ("DMMA", "UBLKCP", "SYNCS", "UTCMMA", "LDTM", "UTMALDG", "DFMA", "BAR", "LDL", "STL"):
                if op.startswith(key):
                    counts[cur][key] += 1


def demangle(n):
    out = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    out = re.sub(r"\(anonymous namespace\)::", "", out)
    return re.sub(r"\(.*", "", out)


print(f"{os.path.basename(lib)}: {len(usage)} kernels (sm_100a)\n")
print(f"{'kernel':62s} {'regs':>5s} {'stack':>6s} {'smem':>6s} {'DMMA':>6s} {'UBLKCP':>7s} {'SYNCS':>6s} {'LDL+STL':>8s} {'tcgen05':>8s}")
tot = collections.Counter()
for n in sorted(usage, key=demangle):
    u, c = usage[n], counts[n]
    print(f"{demangle(n)[:62]:62s} {u.get('REG', '?'):>5s} {u.get('STACK', '?'):>6s} {u.get('SHARED', '?'):>6s} "
          f"{c['DMMA']:6d} {c['UBLKCP']:7d} {c['SYNCS']:6d} {c['LDL'] + c['STL']:8d} {c['UTCMMA'] + c['LDTM'] + c['UTMALDG']:8d}")
    tot.update(c)
rest = [n for n in counts if n not in usage and sum(counts[n].values())]
if rest:
    print("\ndevice functions outside the kernels (the tiles the persistent workers call):")
    for n in sorted(rest, key=demangle):
        c = counts[n]
        print(f"{demangle(n)[:62]:62s} {'':>5s} {'':>6s} {'':>6s} {c['DMMA']:6d} {c['UBLKCP']:7d} {c['SYNCS']:6d} {c['LDL'] + c['STL']:8d} "
              f"{c['UTCMMA'] + c['LDTM'] + c['UTMALDG']:8d}")
        tot.update(c)
print(f"\ntotal: DMMA {tot['DMMA']}, UBLKCP {tot['UBLKCP']}, SYNCS {tot['SYNCS']}, local loads/stores {tot['LDL'] + tot['STL']}, "
      f"UTCMMA/LDTM/UTMALDG {tot['UTCMMA'] + tot['LDTM'] + tot['UTMALDG']} (tcgen05 has no f64 kind: the contractions use DMMA)")
