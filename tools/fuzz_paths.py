#!/usr/bin/env python3
"""Fuzz (dev tool): random robots, scenes (1..64 primitives, random poses), batch sizes; the grid-culled
kernels against the block-cooperative ones, mismatches checked against the oracle's clearance band.
usage: fuzz_paths.py [seconds]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import vamp_mvt_b200 as vmv
from oracle import pyoracle as po
from tests import scenes
from vamp_mvt_b200 import _lib

L = _lib.lib()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(time.time()) % 100000)
t_end = time.time() + budget
n_cases = n_units = n_band = 0
keep = {"panda": 0.0, "ur5": 0.0, "fetch": 0.45, "baxter": 0.5}
while time.time() < t_end:
    robot = ["panda", "ur5", "fetch", "baxter"][int(rng.integers(0, 4))]
    R, O = getattr(vmv, robot), po.Oracle(robot)
    n_obj = int(rng.integers(1, 65))
    a = int(rng.integers(0, n_obj + 1)); b = int(rng.integers(0, n_obj - a + 1))
    span = float(rng.uniform(0.6, 1.6))
    sc = scenes.random_scene(int(rng.integers(0, 1 << 30)), n_spheres=a, n_cuboids=b, n_capsules=n_obj - a - b,
                             lo=(-span, -span, -0.3), hi=(span, span, 1.6), keep_out=keep[robot])
    env = scenes.build_product_env(sc)
    oenv = po.add_scene(po.OracleEnv(), scenes.packed(sc))
    n = int(rng.integers(1024, 20000))
    q = scenes.random_configs(robot, n, seed=int(rng.integers(0, 1 << 30)))
    if rng.random() < 0.3:
        q[: n // 3] += rng.uniform(-0.5, 0.5, size=(n // 3, q.shape[1])).astype(np.float32)
    out = {}
    for path in (2, 3):
        L.vmv_force_kernel_path(path if not (path == 2 and robot == "baxter" and False) else 0)
        out[path] = R.validate_batch(q, env)
    ne = int(rng.integers(64, 3000))
    ea, eb = scenes.random_edges(robot, ne, seed=int(rng.integers(0, 1 << 30)))
    eout = {}
    for path in (1, 3):
        L.vmv_force_kernel_path(path)
        eout[path] = R.validate_motion_batch(ea, eb, env)
    L.vmv_force_kernel_path(0)
    bad = np.nonzero(out[2] != out[3])[0]
    if len(bad):
        clear = O.min_clearance(oenv, q[bad])
        assert np.abs(clear).max() <= 1e-5, (robot, n_obj, n, clear)
        n_band += len(bad)
    ebad = int((eout[1] != eout[3]).sum())
    if ebad:
        want = O.validate_edges(oenv, ea, eb)
        assert (eout[3] != want).sum() <= 2 and (eout[1] != want).sum() <= 2, (robot, n_obj, ne, ebad)
    n_cases += 1
    n_units += n + ne
print(f"fuzz ok: {n_cases} cases, {n_units} units, {n_band} in-band config differences between kernel generations")
