#!/usr/bin/env python3
"""dev tool: build a library variant into variants/lib_<name>.so for A/B timing on the GPU box.
usage: build_variant.py <name> [--units vmv_robot_panda,vmv_host] [nvcc flags ...]
Only the listed translation units (default: all) are recompiled with the extra flags; the others are taken
from build/obj (the default build).  Select a variant at run time with VMV_LIB=variants/lib_<name>.so."""
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import __graft_entry__ as g  # noqa: E402

name = sys.argv[1]
args = sys.argv[2:]
units = list(g.UNITS)
if args and args[0] == "--units":
    units = args[1].split(",")
    args = args[2:]
out_dir = REPO / "variants"
obj_dir = out_dir / f"obj_{name}"
obj_dir.mkdir(parents=True, exist_ok=True)
procs = []
for u in units:
    cmd = ["nvcc", *g.NVCC_FLAGS, *args, "-c", str(g.CSRC / f"{u}.cu"), "-o", str(obj_dir / f"{u}.o")]
    procs.append((u, subprocess.Popen(cmd, cwd=str(g.CSRC), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for u, p in procs:
    out, _ = p.communicate()
    if p.returncode != 0:
        sys.exit(f"nvcc failed on {u}:\n{out}")
    prev = ""
    for line in out.splitlines():
        if "registers" in line and ("v4" in prev or "k_validate" in prev):
            print(u, prev.split("'")[1][:60] if "'" in prev else "", line.strip()[:90])
        prev = line
objs = [str((obj_dir if u in units else g.OBJ_DIR) / f"{u}.o") for u in g.UNITS]
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(out_dir / f"lib_{name}.so"), *objs, "-ldl"], check=True)
print("built", out_dir / f"lib_{name}.so")
