#!/usr/bin/env python3
"""dev tool: random CAPT pointclouds and build parameters, the list-free kernels against the compiled reference (oracle/_ref).
Every mismatching configuration is classified by the oracle's clearance on the same tree (tests/parity.py).
usage: fuzz_capt.py [seconds]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import vamp_mvt_b200 as vmv
from oracle import pyoracle as po
from tests import parity, scenes

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(20260)
t_end = time.time() + budget
n_cases = n_units = n_band = n_invalid = 0
while time.time() < t_end:
    robot = ["panda", "ur5", "fetch"][int(rng.integers(3))]
    R, O, ref = getattr(vmv, robot), po.Oracle(robot), po.Ref(robot)
    n = int(rng.choice([2, 3, 9, 100, 1000, 5000, 20000])) if rng.random() < 0.5 else int(rng.integers(2, 6000))  # (one point: the reference's query crashes)
    kind = int(rng.integers(3))
    if kind == 0:
        pts = rng.uniform([-0.9, -0.9, 0.0], [0.9, 0.9, 1.3], size=(n, 3))
    elif kind == 1:
        pts = np.concatenate([rng.uniform([0.2, -0.7, 0.0], [0.9, 0.7, 0.03], size=(n - n // 3, 3)), rng.normal([0.5, 0.2, 0.5], 0.07, size=(n // 3, 3))])
    else:
        pts = rng.normal([0.45, 0.0, 0.45], [0.25, 0.25, 0.2], size=(n, 3))
    pts = pts.astype(np.float32)
    keep = np.hypot(pts[:, 0], pts[:, 1]) > (0.5 if robot == "fetch" else 0.25)
    pts = pts[keep] if keep.sum() >= 2 else np.float32([[0.6, 0.1, 0.4], [0.7, -0.2, 0.6]])
    # no exact coordinate ties at all (which half a tied point lands in is the reference's unstable sort's choice)
    pts = pts + rng.uniform(-1e-6, 1e-6, size=pts.shape).astype(np.float32)
    r_max = float(rng.choice([0.02, 0.05, 0.08, 0.12, 0.24, 0.4]))
    r_min = float(rng.uniform(0.0, r_max))
    r_point = float(rng.choice([0.0, 0.0025, 0.01, 0.04]))
    env, renv, oenv = vmv.Environment(), po.RefEnv(), po.OracleEnv()
    env.add_capt_pointcloud(pts, r_min, r_max, r_point)
    renv.add_capt(pts, r_min, r_max, r_point)
    oenv.add_capt(pts, r_min, r_max, r_point)
    q = scenes.random_configs(robot, 4096, seed=int(rng.integers(1 << 30)))
    got = R.validate_batch(q, env)
    want = ref.validate_configs(renv, q, threads=8)
    rep = parity.config_report(O, oenv, q, got, want, has_cloud=True)
    assert rep["mismatch_outside"] == 0, (robot, len(pts), r_min, r_max, r_point, rep)
    n_band += rep["mismatch_in_band"]
    n_invalid += int((~want).sum())
    n_cases += 1
    n_units += len(q)
print(f"fuzz_capt ok: {n_cases} clouds, {n_units} configurations ({n_invalid} invalid), {n_band} in-band differences, 0 outside")
