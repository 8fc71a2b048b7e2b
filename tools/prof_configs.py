#!/usr/bin/env python3
"""Small fixed workload for ncu: a few launches of the config kernel (and optionally the edge kernel)
on the bench scene.  Usage: prof_configs.py [configs|edges] [n_log2]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import vamp_mvt_b200 as vmv
from vamp_mvt_b200 import _lib
from tests import scenes

what = sys.argv[1] if len(sys.argv) > 1 else "configs"
nlog2 = int(sys.argv[2]) if len(sys.argv) > 2 else 20
L = _lib.lib()
robot = vmv.panda
if what == "hf":
    # the any-environment kernel without a pointcloud: Fetch vs a 256x256 heightfield
    R = vmv.fetch
    rng = np.random.default_rng(1)
    data = (0.15 * rng.random((256, 256)) ** 4).astype(np.float32)
    env = vmv.Environment()
    env.add_heightfield(vmv.make_heightfield([0, 0, -0.3], [0.02, 0.02, 1.0], [256, 256], data))
    n = 1 << min(nlog2, 18)
    q = scenes.random_configs("fetch", n, seed=0)
    dq = L.vmv_dev_alloc(q.nbytes); db = L.vmv_dev_alloc((n + 31) // 32 * 4)
    _lib.check(L.vmv_memcpy_h2d(dq, _lib.ptr(q), q.nbytes, None))
    for _ in range(3):
        _lib.check(L.vmv_validate_configs_dev(R.id, env.handle, dq, n, db, None))
    _lib.check(L.vmv_stream_sync(None))
elif what == "c4":
    # BASELINE config 4: Fetch vs CAPT of ~10^5 points + heightfield (the any-environment kernel)
    from tests import workloads
    rb = sys.argv[3] if len(sys.argv) > 3 else "fetch"
    R = getattr(vmv, rb)
    env, pts, hf, _ = workloads.c4_environment(rb)
    n = 1 << min(nlog2, 18)
    q = scenes.random_configs(rb, n, seed=0)
    dq = L.vmv_dev_alloc(q.nbytes); db = L.vmv_dev_alloc((n + 31) // 32 * 4)
    _lib.check(L.vmv_memcpy_h2d(dq, _lib.ptr(q), q.nbytes, None))
    for _ in range(3):
        _lib.check(L.vmv_validate_configs_dev(R.id, env.handle, dq, n, db, None))
    _lib.check(L.vmv_stream_sync(None))
elif what == "configs":
    env = scenes.build_product_env(scenes.table_shelf_scene())
    n = 1 << nlog2
    q = scenes.random_configs("panda", n, seed=0)
    dq = L.vmv_dev_alloc(q.nbytes); db = L.vmv_dev_alloc((n + 31) // 32 * 4)
    _lib.check(L.vmv_memcpy_h2d(dq, _lib.ptr(q), q.nbytes, None))
    for _ in range(4):
        _lib.check(L.vmv_validate_configs_dev(robot.id, env.handle, dq, n, db, None))
    _lib.check(L.vmv_stream_sync(None))
else:
    env = scenes.build_product_env(scenes.box_scene())
    n = 1 << min(nlog2, 18)
    a, b = scenes.random_edges("panda", n, seed=0)
    da = L.vmv_dev_alloc(a.nbytes); dbb = L.vmv_dev_alloc(b.nbytes); db = L.vmv_dev_alloc((n + 31) // 32 * 4)
    _lib.check(L.vmv_memcpy_h2d(da, _lib.ptr(a), a.nbytes, None)); _lib.check(L.vmv_memcpy_h2d(dbb, _lib.ptr(b), b.nbytes, None))
    for _ in range(4):
        _lib.check(L.vmv_validate_edges_dev(robot.id, env.handle, da, dbb, n, 0, db, None))
    _lib.check(L.vmv_stream_sync(None))
print("ok")
