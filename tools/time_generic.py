#!/usr/bin/env python3
"""Any-environment kernel timing (dev tool): 2^18 configurations per robot against a 256x256 heightfield
(no pointcloud: no scan tail), device pointers.  VMV_LIB selects the library build."""
import json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import vamp_mvt_b200 as vmv
from vamp_mvt_b200 import _lib
from tests import scenes

L = _lib.lib()
rng = np.random.default_rng(1)
xd = yd = 256
data = (0.15 * rng.random((yd, xd)) ** 4).astype(np.float32)
out = {"lib": os.environ.get("VMV_LIB", "default")}
for robot in ("panda", "ur5", "fetch", "baxter"):
    R = getattr(vmv, robot)
    env = vmv.Environment()
    env.add_heightfield(vmv.make_heightfield([0, 0, -0.3], [0.02, 0.02, 1.0], [xd, yd], data))
    n = 1 << 18
    q = scenes.random_configs(robot, n, seed=0)
    dq, db = L.vmv_dev_alloc(q.nbytes), L.vmv_dev_alloc((n + 31) // 32 * 4)
    _lib.check(L.vmv_memcpy_h2d(dq, _lib.ptr(q), q.nbytes, None))
    fn = lambda: _lib.check(L.vmv_validate_configs_dev(R.id, env.handle, dq, n, db, None))
    for _ in range(2):
        fn()
    _lib.check(L.vmv_stream_sync(None))
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    _lib.check(L.vmv_stream_sync(None))
    t = (time.perf_counter() - t0) / 5
    words = np.zeros((n + 31) // 32, np.uint32)
    _lib.check(L.vmv_memcpy_d2h(_lib.ptr(words), db, words.nbytes, None))
    _lib.check(L.vmv_stream_sync(None))
    out[robot] = {"configs_per_s": n / t, "ms": t * 1e3, "valid": float(_lib.unpack_bits(words, n).mean())}
print(json.dumps(out))
