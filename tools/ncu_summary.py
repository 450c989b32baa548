"""Summarise an .ncu-rep (read here, no GPU) into the handful of metrics DESIGN.md quotes.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__grid_size",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct",
        "smsp__average_warps_issue_stalled", "smsp__warps_eligible.avg", "lts__t_bytes.sum", "sm__inst_executed_pipe_lsu",
        "smsp__inst_executed.sum"]
SKIP = ("per_second", "bytes_read.sum.pct", "bytes_write.sum.pct")

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = [i for i, h in enumerate(hdr) if any(w in h for w in WANT) and not any(s in h for s in SKIP)]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print("=====", r[ki])
    for i in idx:
        if r[i] not in ("0", "", "0.000000") or "stalled" not in hdr[i]:
            print(f"  {hdr[i]} = {r[i]} {units[i]}")
