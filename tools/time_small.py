#!/usr/bin/env python3
"""Latency of one device-pointer call at small batch sizes, per kernel generation (dev tool: where the
automatic choice between the block-cooperative and the grid-culled kernels should switch)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import vamp_mvt_b200 as vmv
from vamp_mvt_b200 import _lib
from tests import scenes

L = _lib.lib()
env = scenes.build_product_env(scenes.table_shelf_scene())
h = env.handle
q = scenes.random_configs("panda", 1 << 16, seed=0)
a, b = scenes.random_edges("panda", 1 << 14, seed=0)
dq, da, db_ = L.vmv_dev_alloc(q.nbytes), L.vmv_dev_alloc(a.nbytes), L.vmv_dev_alloc(b.nbytes)
bits = L.vmv_dev_alloc(1 << 16)
for d, x in ((dq, q), (da, a), (db_, b)):
    _lib.check(L.vmv_memcpy_h2d(d, _lib.ptr(x), x.nbytes, None))
for what, sizes in (("configs", [32, 128, 512, 2048, 4096, 16384]), ("edges", [32, 64, 128, 256, 1024, 4096])):
    for n in sizes:
        row = []
        for path in (1, 2, 3):
            L.vmv_force_kernel_path(path)
            call = (lambda: L.vmv_validate_configs_dev(vmv.panda.id, h, dq, n, bits, None)) if what == "configs" else \
                   (lambda: L.vmv_validate_edges_dev(vmv.panda.id, h, da, db_, n, 0, bits, None))
            for _ in range(5):
                _lib.check(call())
            _lib.check(L.vmv_stream_sync(None))
            t0 = time.perf_counter()
            for _ in range(200):
                call()
            _lib.check(L.vmv_stream_sync(None))
            row.append((time.perf_counter() - t0) / 200 * 1e6)
        L.vmv_force_kernel_path(0)
        print(f"{what:8s} n={n:6d}: per-thread {row[0]:8.1f} us   block {row[1]:8.1f} us   grid {row[2]:8.1f} us")
