"""Warp-stall samples of an .ncu-rep (captured with --set full --import-source on, library built with
-lineinfo) aggregated by source line and by region of vl_kernels.cu, plus the headline metrics.
    python tools/ncu_hotspots.py report.ncu-rep > profiles/<name>.txt
"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
WANT = ("gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed")
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=====", d["Kernel Name"][:100])
    for k in WANT:
        if k in d:
            print(f"  {k} = {d[k]} {rows[1][hdr.index(k)]}")
    for k in hdr:
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
            try:
                if float(d[k]) >= 0.3:
                    print(f"  {k} = {d[k]}")
            except ValueError:
                pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, h, data = None, None, []
for r in csv.reader(src.splitlines()):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 5 and r[0] == "Line No":
        h = r
    elif h and len(r) == len(h) and r[0] != "":
        try:
            data.append((int(r[4]), cur, int(r[0]), r[1], dict(zip(h, r))))
        except ValueError:
            pass
tot = sum(x[0] for x in data) or 1
text = open(__file__.replace("tools/ncu_hotspots.py", "viloca_b200/csrc/vl_kernels.cu")).read().splitlines()


def region(f, ln):
    if f != "vl_kernels.cu":
        return f
    name = "?"
    for i in range(min(ln, len(text)) - 1, -1, -1):        # nearest preceding function definition
        t = text[i]
        for key in ("vl_sched_claim", "vl_wait_published", "vl_split_combine", "vl_haplo_body", "vl_cluster_tile_mt", "vl_cluster_close",
                    "vl_sched_kernel", "vl_sched_persistent_kernel", "vl_cluster_body"):
            if key + "(" in t and ("__device__" in t or "__global__" in text[max(i - 1, 0)] or "__device__" in text[max(i - 1, 0)] or t.startswith(key)):
                return key
    return name


agg = collections.Counter()
for s, f, ln, _, _ in data:
    agg[region(f, ln)] += s
print(f"\nwarp-stall samples by function ({tot} samples)")
for k, v in agg.most_common():
    print(f"  {100 * v / tot:5.1f}%  {k}")
keys = [k for k in (h or []) if k.startswith("stall_") and "Not Issued" not in k]
print("\nhottest source lines")
for s, f, ln, line, d in sorted(data, key=lambda x: -x[0])[:30]:
    top = sorted(((int(d[k] or 0), k) for k in keys), reverse=True)[:2]
    print(f"  {100 * s / tot:5.1f}% {f}:{ln:<5d} {line.strip()[:80]:80s} | {top[0][1]}={top[0][0]} {top[1][1]}={top[1][0]}")
