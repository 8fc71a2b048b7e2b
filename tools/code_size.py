#!/usr/bin/env python3
"""dev tool: static code size of one kernel by source file / line range, from `nvdisasm -g -c` of the unit's cubin.
usage: code_size.py <unit .o> <kernel name substring> [file:lo:hi:name ...]"""
import collections, re, subprocess, sys, tempfile, os

obj, pat = sys.argv[1], sys.argv[2]
ranges = [a.split(":") for a in sys.argv[3:]]
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    text = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
cur, on = None, False
by_file, by_line = collections.Counter(), collections.Counter()
total = 0
for line in text.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", line)
    if m:
        on = pat in m.group(1)
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and cur:
        total += 1
        by_file[cur[0]] += 1
        by_line[cur] += 1
print(f"{total} instructions = {total * 16 / 1024:.1f} KB")
for f, n in by_file.most_common():
    print(f"  {f:32s} {n * 16 / 1024:7.1f} KB")
for f, lo, hi, name in ranges:
    n = sum(v for (ff, l), v in by_line.items() if ff == f and int(lo) <= l <= int(hi))
    print(f"  range {name:20s} {n * 16 / 1024:7.1f} KB")
if not ranges:
    print("largest lines:")
    for (f, l), n in by_line.most_common(25):
        print(f"  {f}:{l:<5d} {n * 16:6d} B")
