#!/usr/bin/env python3
"""Small fixed workload for compute-sanitizer: every kernel generation on a few thousand units."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import vamp_mvt_b200 as vmv
from vamp_mvt_b200 import _lib
from tests import scenes

L = _lib.lib()
env = scenes.build_product_env(scenes.table_shelf_scene())
q = scenes.random_configs("panda", 4096 + 17, seed=0)
a, b = scenes.random_edges("panda", 300, seed=0)
ref = None
for path in (1, 2, 3):
    L.vmv_force_kernel_path(path)
    v = vmv.panda.validate_batch(q, env)
    e = vmv.panda.validate_motion_batch(a, b, env)
    if ref is None:
        ref = (v, e)
    print("path", path, "valid", v.mean(), e.mean(), "agree", (v == ref[0]).all(), (e == ref[1]).all())
L.vmv_force_kernel_path(0)
rng = np.random.default_rng(0)
pts = rng.uniform([0.3, -0.6, 0.0], [0.9, 0.6, 0.4], size=(3000, 3)).astype(np.float32)
_, keep = np.unique(np.floor(pts / 0.025).astype(np.int64), axis=0, return_index=True)
pts = pts[np.sort(keep)]
env2 = vmv.Environment()
env2.add_capt_pointcloud(pts, 0.012, 0.08, vmv.POINT_RADIUS)
env2.add_mvt_pointcloud(pts, 0.012, 0.08, [-1.5, -1.5, -0.5], [1.5, 1.5, 2.5], vmv.POINT_RADIUS)
print("clouds", vmv.panda.validate_batch(q[:2000], env2).mean(), vmv.panda.validate_motion_batch(a[:100], b[:100], env2).mean())
for r in ("ur5", "fetch", "baxter"):
    R = getattr(vmv, r)
    sc = scenes.random_scene(1, keep_out=0.5)
    print(r, R.validate_batch(scenes.random_configs(r, 4200, seed=1), scenes.build_product_env(sc)).mean())
print("ok")
# the list-free CAPT on the warp-autonomous any-environment kernel (>= 1024 configurations), device-built trees of several sizes
for n in (1, 2, 5, 700, 3000):
    envc = vmv.Environment()
    envc.add_capt_pointcloud(pts[:n], 0.012, 0.08, vmv.POINT_RADIUS)
    print("capt", n, vmv.panda.validate_batch(q[:2048], envc).mean())
hf = vmv.make_heightfield([0, 0, -0.3], [0.05, 0.05, 1.0], [32, 32], (0.1 * rng.random((32, 32))).astype(np.float32))
envh = vmv.Environment()
envh.add_capt_pointcloud(pts, 0.03, 0.24, vmv.POINT_RADIUS)
envh.add_heightfield(hf)
for r in ("panda", "ur5", "fetch"):
    R = getattr(vmv, r)
    print("capt + heightfield", r, R.validate_batch(scenes.random_configs(r, 2048, seed=2), envh).mean())
print("ok 2")
