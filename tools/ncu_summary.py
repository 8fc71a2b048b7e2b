#!/usr/bin/env python3
"""Text summary of an .ncu-rep for profiles/: headline details metrics, DRAM bytes, pipe activity,
and the per-file / per-line instruction and stall-sample shares (needs -lineinfo + --import-source).
Usage: ncu_summary.py <rep> > profiles/<name>_ncu_summary.txt"""
import csv, io, subprocess, sys

rep = sys.argv[1]
KEYS = ("Duration", "Executed Ipc Active", "Executed Ipc Elapsed", "Issue Slots Busy", "No Eligible", "Memory Throughput", "DRAM Throughput",
        "L1/TEX Hit Rate", "L2 Hit Rate", "Compute (SM) Throughput", "Active Warps Per Scheduler", "Eligible Warps Per Scheduler",
        "Avg. Active Threads Per Warp", "Registers Per Thread", "Dynamic Shared Memory Per Block", "Block Size", "Grid Size",
        "Block Limit", "Theoretical Occupancy", "Achieved Occupancy", "Achieved Active Warps Per SM")
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
for line in det.splitlines():
    t = line.strip()
    if t.startswith("void ") or t.startswith("k_") or any(t.startswith(k) for k in KEYS):
        print(line.rstrip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
if len(rows) >= 3:
    h, u, v = rows[0], rows[1], rows[2]
    print("  raw metrics:")
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "smsp__inst_executed.sum",
              "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
              "lts__t_sectors.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active"):
        if k in h:
            i = h.index(k)
            print(f"    {k:70s} {v[i]} {u[i]}")
print("  source-level shares (warp instructions executed / stall samples):")
out = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_phase_breakdown.py"), rep, "14"], capture_output=True, text=True).stdout
for line in out.splitlines():
    print("  " + line)
