#!/usr/bin/env python3
"""dev tool: edge batches against BASELINE config 4's environment (CAPT + heightfield), device pointers, CUDA events."""
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
N = 1 << 16
for rb in sys.argv[1:] or ["fetch", "ur5"]:
    R = getattr(vmv, rb)
    env, pts, hf, _ = workloads.c4_environment(rb)
    a, b = scenes.random_edges(rb, N, seed=0)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    bits = torch.zeros((N + 31) // 32, dtype=torch.int32, device="cuda")
    run = lambda: _lib.check(L.vmv_validate_edges_dev(R.id, env.handle, da.data_ptr(), db.data_ptr(), N, 0, bits.data_ptr(), stream))
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 5
    v = float(np.unpackbits(bits.cpu().numpy().view(np.uint8)).mean())
    print(f"{os.environ.get('VMV_LIB', 'default'):24s} c4 edges {rb:6s} {t:.3f} ms  {N / t / 1e3:.2f} M edges/s  valid {v:.4f}")
