#!/usr/bin/env python3
"""Host-buffer path timing (dev tool): raw pinned H2D bandwidth next to vmv_validate_configs."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import vamp_mvt_b200 as vmv
from vamp_mvt_b200 import _lib
from tests import scenes

L = _lib.lib()
N = 1 << 20
env = scenes.build_product_env(scenes.table_shelf_scene())
h = env.handle
qs = [torch.from_numpy(scenes.random_configs("panda", N, seed=b)).pin_memory() for b in range(4)]
bits = torch.zeros((N + 31) // 32, dtype=torch.int32).pin_memory()
d = torch.empty_like(qs[0], device="cuda")
for _ in range(3):
    d.copy_(qs[0], non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20):
    d.copy_(qs[i % 4], non_blocking=True)
torch.cuda.synchronize()
t_copy = (time.perf_counter() - t0) / 20
for i in range(3):
    _lib.check(L.vmv_validate_configs(vmv.panda.id, h, qs[i % 4].data_ptr(), N, bits.data_ptr()))
t0 = time.perf_counter()
for i in range(20):
    _lib.check(L.vmv_validate_configs(vmv.panda.id, h, qs[i % 4].data_ptr(), N, bits.data_ptr()))
t_e2e = (time.perf_counter() - t0) / 20
print(f"{os.environ.get('VMV_LIB', 'default'):36s} chunk_log2={os.environ.get('VMV_CHUNK_LOG2', '17')}: raw H2D {t_copy * 1e3:.3f} ms ({qs[0].numel() * 4 / t_copy / 1e9:.1f} GB/s); "
      f"e2e {t_e2e * 1e3:.3f} ms ({N / t_e2e / 1e9:.2f} G configs/s)")
