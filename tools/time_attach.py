#!/usr/bin/env python3
"""dev tool: configuration and edge batches with an ATTACHMENT on a primitive scene: the grid-culled kernels with phase D against
the per-thread kernels (VMV_NO_V4_ATTACH=1).  Prints one line per robot."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import vamp_mvt_b200 as vmv
from vamp_mvt_b200 import _lib
from tests import scenes

L = _lib.lib()
stream = torch.cuda.current_stream().cuda_stream
N, NE = 1 << 20, 1 << 16
for rb in sys.argv[1:] or ["panda", "fetch"]:
    R = getattr(vmv, rb)
    sc = scenes.table_shelf_scene() if rb == "panda" else scenes.random_scene(2, keep_out={"ur5": 0.0, "fetch": 0.45, "baxter": 0.5}[rb])
    env = scenes.build_product_env(sc)
    tf = np.eye(4, dtype=np.float32)
    tf[:3, 3] = [0.0, 0.0, 0.08]
    att = vmv.Attachment(tf)
    att.add_spheres([vmv.Sphere([0, 0, 0], 0.04), vmv.Sphere([0, 0, 0.06], 0.03), vmv.Sphere([0.03, 0, 0.1], 0.025)])
    env.attach(att)
    q = torch.from_numpy(scenes.random_configs(rb, N, seed=0)).cuda()
    a, b = scenes.random_edges(rb, NE, seed=0)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    bits = torch.zeros((N + 31) // 32, dtype=torch.int32, device="cuda")

    def timeit(fn, reps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    tc = timeit(lambda: _lib.check(L.vmv_validate_configs_dev(R.id, env.handle, q.data_ptr(), N, bits.data_ptr(), stream)), 5)
    vc = float(np.unpackbits(bits.cpu().numpy().view(np.uint8)).mean())
    te = timeit(lambda: _lib.check(L.vmv_validate_edges_dev(R.id, env.handle, da.data_ptr(), db.data_ptr(), NE, 0, bits.data_ptr(), stream)), 3)
    print(f"{'per-thread' if os.environ.get('VMV_NO_V4_ATTACH') else 'grid-culled'} {rb:6s} attached: configs {tc:.3f} ms ({N / tc / 1e3:.0f} M/s, valid {vc:.4f})  "
          f"edges {te:.3f} ms ({NE / te / 1e3:.1f} M/s)")
