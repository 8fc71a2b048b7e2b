#!/usr/bin/env python3
"""Generates tests/golden/frontends.npz from the reference's own code compiled in place
(oracle/_ref/libvamp_ref.so): the Halton sampler, Path members, simplify and the CenterVox filter.
Run where /root/reference is mounted; the fixture is committed, this script documents how it was made.
The cases are the ones tests/test_simplify.py and tests/test_pointcloud_filter.py run live against the
compiled reference; the fixture lets them be checked where that library is not available."""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import numpy as np

from oracle import pyoracle as po
from tests import scenes
from tests import test_pointcloud_filter as tf
from tests import test_simplify as ts

out = {}
for robot in ("panda", "ur5", "fetch", "baxter"):
    ref = po.Ref(robot)
    out[f"halton_{robot}"] = ref.halton(512)
    out[f"halton_{robot}_at_999990"] = ref.halton(32, skip=999990)
    wp = scenes.random_configs(robot, 6, seed=77)
    out[f"path_{robot}_in"] = wp
    for op, arg in ((1, 0), (2, 8), (3, 23)):
        p, cost = ref.path_op(op, wp, arg)
        out[f"path_{robot}_op{op}"] = p
        out[f"path_{robot}_op{op}_cost"] = np.float32(cost)

cases = [0, 1, 2, 4, 7, 9, 10]
out["simplify_cases"] = np.array(cases, np.int32)
for c in cases:
    robot, scene_name, ops, opts = ts.CASES[c]
    ref = po.Ref(robot)
    renv = po.add_scene(po.RefEnv(), scenes.packed(ts.scene_of(scene_name, robot)))
    path = ts.jagged_path(robot, ref, renv, 0)
    st = ts.make_settings(opts, ops)
    samples = np.random.default_rng(100).random((257, ref.dof), dtype=np.float32)
    want, it = ref.simplify(renv, path, ops, ts.settings12(st), samples)
    out[f"simplify_{c}_in"], out[f"simplify_{c}_samples"], out[f"simplify_{c}_out"] = path, samples, want
    out[f"simplify_{c}_iterations"] = np.int32(it)

fcases = [1, 3, 4, 5]
out["filter_cases"] = np.array(fcases, np.int32)
for c in fcases:
    kind, n, vox, rng_, origin, lo, hi = tf.CASES[c]
    p = tf.with_specials(tf.cloud(kind, min(n, 20000), 1), 1)
    out[f"filter_{c}_in"] = p
    out[f"filter_{c}_out"] = po.ref_filter_centervox(p, vox, rng_, origin, lo, hi)

np.savez_compressed(REPO / "tests" / "golden" / "frontends.npz", **out)
print("wrote", REPO / "tests" / "golden" / "frontends.npz", len(out), "arrays")
