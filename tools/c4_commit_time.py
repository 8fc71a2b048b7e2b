#!/usr/bin/env python3
"""Commit time of the C4 Fetch environment (dev tool): CAPT of ~10^5 points -- upload of the afford lists + the
clearance grid build.  VMV_LIB selects the library build."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
import vamp_mvt_b200 as vmv
from vamp_mvt_b200 import _lib
from tools import bench_extra as be
L = _lib.lib()
R = vmv.fetch
rmin, rmax = R.min_max_radii()
pts = be.synth_pointcloud(100_000, 0.55)
env = vmv.Environment()
env.add_capt_pointcloud(pts, rmin, rmax, vmv.POINT_RADIUS)
t0 = time.perf_counter(); h = env.handle; _lib.check(L.vmv_stream_sync(None)); print("commit s", time.perf_counter() - t0)
