#!/usr/bin/env python3
"""dev tool: query statistics of the any-environment kernel on BASELINE config 4 (needs a -DVMV_C4_STATS build:
tools/build_variant.py stats --units vmv_robot_fetch,vmv_robot_ur5 -DVMV_C4_STATS; VMV_LIB=variants/lib_stats.so)."""
import ctypes as C, json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import vamp_mvt_b200 as vmv
from vamp_mvt_b200 import _lib
from tests import scenes, workloads

L = _lib.lib()
L.vmv_dev_stats.argtypes = [C.c_void_p, C.c_int]
NAMES = ["sphere_queries", "hit_primitives", "grid_lookups", "undecided_after_grid", "capt_active_calls", "need_enumeration", "enumerations",
         "enumeration_steps", "points_loaded", "enumeration_hits", "group_size_sum", "warp_calls", "table_proven_hits", "table_point_not_listed", "bounding_hits", "fine_hits"]
for robot in sys.argv[1:] or ["fetch", "ur5"]:
    R = getattr(vmv, robot)
    env, pts, hf, _ = workloads.c4_environment(robot)
    n = 1 << 16
    q = scenes.random_configs(robot, n, seed=0)
    out = np.zeros(64, np.uint64)
    R.validate_batch(q[:64], env)
    L.vmv_dev_stats(None, 1)
    v = R.validate_batch(q, env)
    L.vmv_dev_stats(_lib.ptr(out), 1)
    d = {k: float(out[i]) / n for i, k in enumerate(NAMES)}
    d["valid"] = float(v.mean())
    d["points_read_hist_hits(32,64,..)"] = [int(x) for x in out[16:28]]
    d["points_read_hist_nohits"] = [int(x) for x in out[28:40]]
    d["scans_by_query_radius(<.05,<.1,<.2,>=.2)"] = [int(x) for x in out[40:44]]
    d["nohit_scans_by_query_radius"] = [int(x) for x in out[44:48]]
    d["tile_cycles_hist(2^13,2^14,..)"] = [int(x) for x in out[48:60]]
    d["cycles_per_tile: capt call / of which scan loop / probe"] = [float(out[k]) / max(1.0, float(out[61])) for k in (14, 15, 62)]
    d["enum: lookup rounds / cells with points / cells in box (per config)"] = [float(out[k]) / n for k in (16, 17, 18)]
    d["index_violations (must be 0)"] = int(out[63])
    d["tile_cycles_mean"] = float(out[60]) / max(1.0, float(out[61]))
    print(robot, json.dumps(d, indent=1))
