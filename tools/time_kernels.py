#!/usr/bin/env python3
"""Kernel timing for A/B runs (dev tool): C2 configs and C3 edges, device pointers, CUDA events.
VMV_LIB selects the library build.  Prints one line."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import vamp_mvt_b200 as vmv
from vamp_mvt_b200 import _lib
from tests import scenes

L = _lib.lib()
robot = vmv.panda
stream = torch.cuda.current_stream().cuda_stream
N, NE, NB = 1 << 20, 1 << 18, 8
env = scenes.build_product_env(scenes.table_shelf_scene())
batches = [torch.from_numpy(scenes.random_configs("panda", N, seed=b)).cuda() for b in range(NB)]
bits = torch.zeros((N + 31) // 32, dtype=torch.int32, device="cuda")
box = scenes.build_product_env(scenes.box_scene())
a, b = scenes.random_edges("panda", NE, seed=0)
a_d, b_d = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
ebits = torch.zeros((NE + 31) // 32, dtype=torch.int32, device="cuda")

def timeit(fn, reps):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

tc = timeit(lambda i: _lib.check(L.vmv_validate_configs_dev(robot.id, env.handle, batches[i % NB].data_ptr(), N, bits.data_ptr(), stream)), 40)
vc = float(np.unpackbits(bits.cpu().numpy().view(np.uint8)).mean())
te = timeit(lambda i: _lib.check(L.vmv_validate_edges_dev(robot.id, box.handle, a_d.data_ptr(), b_d.data_ptr(), NE, 0, ebits.data_ptr(), stream)), 10)
ve = float(np.unpackbits(ebits.cpu().numpy().view(np.uint8)).mean())
print(f"{os.environ.get('VMV_LIB', 'default'):40s} configs {tc:.4f} ms ({N / tc / 1e3:.0f} M/s, valid {vc:.5f})  edges {te:.4f} ms ({NE / te / 1e3:.1f} M/s, valid {ve:.5f})")
