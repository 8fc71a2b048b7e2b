#!/usr/bin/env python3
"""Sampling front-end timing (dev tool): 10^6 Halton samples of the Panda vs the C2 scene,
(a) generated on the host path's way: configurations uploaded through vmv_validate_configs (pinned),
(b) generated and validated on the device: vmv_validate_halton (one bit per sample back),
(c) the fill kernel alone (CUDA events).  Prints one JSON line."""
import json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import vamp_mvt_b200 as vmv
from vamp_mvt_b200 import _lib
from tests import scenes

L = _lib.lib()
R = vmv.panda
N = 1000000
env = scenes.build_product_env(scenes.table_shelf_scene())
q = R.halton().take(N)
pinned = torch.from_numpy(q).pin_memory()
words = torch.zeros((N + 31) // 32, dtype=torch.int32).pin_memory()


def wall(fn, reps=20):
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


t_up = wall(lambda: _lib.check(L.vmv_validate_configs(R.id, env.handle, pinned.data_ptr(), N, words.data_ptr())))
bits_up = words.numpy().copy()
t_dev = wall(lambda: _lib.check(L.vmv_validate_halton(R.id, env.handle, 0, N, words.data_ptr(), None)))
assert np.array_equal(bits_up, words.numpy())
d_q = torch.empty((N, 7), dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(L.vmv_halton_fill_dev(R.id, 0, N, d_q.data_ptr(), stream))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    _lib.check(L.vmv_halton_fill_dev(R.id, 0, N, d_q.data_ptr(), stream))
e1.record()
torch.cuda.synchronize()
t_fill = e0.elapsed_time(e1) / 50
assert np.array_equal(d_q.cpu().numpy(), q)
print(json.dumps({"tool": "time_halton", "n": N, "upload_path_ms": round(t_up, 4), "upload_path_configs_per_s": N / t_up * 1e3,
                  "device_sampler_ms": round(t_dev, 4), "device_sampler_configs_per_s": N / t_dev * 1e3,
                  "fill_kernel_ms": round(t_fill, 5), "fill_kernel_GBps": N * 28 / t_fill / 1e6}))
