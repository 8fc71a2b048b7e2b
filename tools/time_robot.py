#!/usr/bin/env python3
"""Kernel timing of one robot on the 15-primitive random scene (dev tool for A/B of library variants).
usage: time_robot.py <robot>   (VMV_LIB selects the build; a VMV_DEV_ROBOT build answers for any id
with that robot's kernels, so pass the matching robot)"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import vamp_mvt_b200 as vmv
from vamp_mvt_b200 import _lib
from tests import scenes

robot = sys.argv[1]
keep = {"panda": 0.0, "ur5": 0.0, "fetch": 0.45, "baxter": 0.5}[robot]
L = _lib.lib()
R = getattr(vmv, robot)
stream = torch.cuda.current_stream().cuda_stream
env = scenes.build_product_env(scenes.random_scene(2, keep_out=keep))
N = 1 << 20
qs = [torch.from_numpy(scenes.random_configs(robot, N, seed=b)).cuda() for b in range(4)]
bits = torch.zeros((N + 31) // 32, dtype=torch.int32, device="cuda")
for i in range(3):
    _lib.check(L.vmv_validate_configs_dev(R.id, env.handle, qs[i % 4].data_ptr(), N, bits.data_ptr(), stream))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10):
    _lib.check(L.vmv_validate_configs_dev(R.id, env.handle, qs[i % 4].data_ptr(), N, bits.data_ptr(), stream))
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10
print(f"{os.environ.get('VMV_LIB', 'default'):40s} {robot}: {t:.4f} ms ({N / t / 1e3:.0f} M configs/s), valid {np.unpackbits(bits.cpu().numpy().view(np.uint8)).mean():.5f}")
