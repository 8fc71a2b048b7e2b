#!/usr/bin/env python3
"""dev tool: CAPT environment build time, device build vs host build (VMV_CAPT_HOST_BUILD), the C4 clouds."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import vamp_mvt_b200 as vmv
from tests import workloads

for rb in ("fetch", "ur5"):
    R = getattr(vmv, rb)
    rmin, rmax = R.min_max_radii()
    pts = workloads.synth_pointcloud(100_000, workloads.C4_KEEP_OUT[rb])
    for host in (False, True, False, True):
        if host:
            os.environ["VMV_CAPT_HOST_BUILD"] = "1"
        else:
            os.environ.pop("VMV_CAPT_HOST_BUILD", None)
        env = vmv.Environment()
        t0 = time.perf_counter()
        env.add_capt_pointcloud(pts, rmin, rmax, vmv.POINT_RADIUS)
        t1 = time.perf_counter()
        R.validate_batch(np.zeros((64, R.dimension()), np.float32), env)  # commit: uploads + nearest-point table
        t2 = time.perf_counter()
        print(f"{rb:6s} {'host  ' if host else 'device'} build {1e3 * (t1 - t0):8.2f} ms   commit + first batch {1e3 * (t2 - t1):8.2f} ms")
