#!/usr/bin/env python3
"""Per-source-line instruction / stall-sample breakdown of one kernel from an .ncu-rep captured with
--import-source on (-lineinfo build).  Usage: ncu_phase_breakdown.py <rep> [top_n]"""
import csv, subprocess, sys, collections, io

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr = None, None
agg = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        i_inst = hdr.index("Instructions Executed")
        i_samp = hdr.index("# Samples")
        i_thr = hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or r[0] == "":
        continue
    try:
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, [0, 0, 0, r[1]])
        a[0] += int(r[i_inst]); a[1] += int(r[i_samp]); a[2] += int(r[i_thr])
    except (ValueError, IndexError):
        pass
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values()); tot_t = sum(a[2] for a in agg.values())
print(f"total warp-inst {tot_i}  samples {tot_s}  thread-inst {tot_t}")
byfile = collections.Counter(); sfile = collections.Counter()
for (f, l), a in agg.items():
    byfile[f] += a[0]; sfile[f] += a[1]
for f in byfile:
    print(f"  {f:28s} inst {100*byfile[f]/tot_i:5.1f}%  samples {100*sfile[f]/max(tot_s,1):5.1f}%")
print("top lines by samples:")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"  {f}:{l:<5d} inst {100*a[0]/tot_i:5.1f}%  samp {100*a[1]/max(tot_s,1):5.1f}%  thr/inst {a[2]/max(a[0],1):4.1f}  {a[3].strip()[:90]}")

# optional phase ranges for vmv_kernels_v2.cuh
if len(sys.argv) > 3:
    ranges = []
    for spec in sys.argv[3].split(","):
        name, lo, hi = spec.split(":")
        ranges.append((name, int(lo), int(hi)))
    ph = collections.OrderedDict()
    for (f, l), a in agg.items():
        name = f
        if f == (sys.argv[4] if len(sys.argv) > 4 else "vmv_kernels_v2.cuh"):
            name = "v2:other"
            for n, lo, hi in ranges:
                if lo <= l <= hi:
                    name = "v2:" + n
        p = ph.setdefault(name, [0, 0, 0])
        p[0] += a[0]; p[1] += a[1]; p[2] += a[2]
    print("phases:")
    for n, p in ph.items():
        print(f"  {n:30s} inst {100*p[0]/tot_i:5.1f}%  samp {100*p[1]/max(tot_s,1):5.1f}%  thr/inst {p[2]/max(p[0],1):4.1f}")
