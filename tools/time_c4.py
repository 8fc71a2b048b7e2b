#!/usr/bin/env python3
"""Kernel timing for A/B runs (dev tool): BASELINE config 4 (CAPT + heightfield), Fetch and UR5, device pointers,
CUDA events.  VMV_LIB selects the library build.  Prints one line per robot."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import vamp_mvt_b200 as vmv
from vamp_mvt_b200 import _lib
from tests import scenes, workloads

L = _lib.lib()
stream = torch.cuda.current_stream().cuda_stream
N, NB = 1 << 18, 4
for rb in sys.argv[1:] or ["fetch", "ur5"]:
    R = getattr(vmv, rb)
    t0 = time.time()
    env, pts, hf, _ = workloads.c4_environment(rb)
    t1 = time.time()
    batches = [torch.from_numpy(scenes.random_configs(rb, N, seed=b)).cuda() for b in range(NB)]
    bits = torch.zeros((N + 31) // 32, dtype=torch.int32, device="cuda")

    def run(i):
        _lib.check(L.vmv_validate_configs_dev(R.id, env.handle, batches[i % NB].data_ptr(), N, bits.data_ptr(), stream))

    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    t2 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20
    v = float(np.unpackbits(bits.cpu().numpy().view(np.uint8)).mean())
    print(f"{os.environ.get('VMV_LIB', 'default'):32s} c4 {rb:6s} {t:.4f} ms  {N / t / 1e3:.1f} M configs/s  valid {v:.5f}  (host build {t1 - t0:.1f} s, commit+warm {t2 - t1:.1f} s)")
