"""``vamp.filter_pointcloud`` (reference bindings/environment.cc:212-239).

``filter_type="centervox"`` -- the fork's CenterVox voxel filter (collision/filter_centervox.hh) -- runs on the
GPU (``vmv_filter_pointcloud_centervox``: one atomicMin per point into a voxel key table) and returns the
reference's points in the reference's order.  ``filter_type="scdf"`` (collision/filter.hh:176-273: a greedy
scan along six Morton orders over an unstable sort) is inherently sequential and not provided."""
from __future__ import annotations

import ctypes as C
import time
from typing import Tuple

import numpy as np

from . import _lib


def filter_pointcloud_centervox(pc, voxel_size: float, max_range: float, origin, workcell_min, workcell_max,
                                return_indices: bool = False) -> np.ndarray:
    p = _lib.f32(pc).reshape(-1, 3)
    idx = np.zeros(32768, np.uint32)
    k = C.c_size_t(0)
    L = _lib.lib()
    _lib.check(L.vmv_filter_pointcloud_centervox(
        _lib.ptr(p), len(p), float(voxel_size), float(max_range), _lib.ptr(_lib.f32(origin).reshape(3)),
        _lib.ptr(_lib.f32(workcell_min).reshape(3)), _lib.ptr(_lib.f32(workcell_max).reshape(3)), _lib.ptr(idx), len(idx), C.byref(k)))
    idx = idx[: k.value].astype(np.int64)
    return idx if return_indices else p[idx]


def filter_pointcloud(pc, min_dist: float, max_range: float, voxel_size: float, origin, workcell_min, workcell_max,
                      cull: bool = True, filter_type: str = "centervox") -> Tuple[np.ndarray, int]:
    """Same arguments and return value as the reference's ``vamp.filter_pointcloud``: (points, nanoseconds)."""
    if filter_type == "centervox":
        t0 = time.perf_counter_ns()
        out = filter_pointcloud_centervox(pc, voxel_size, max_range, origin, workcell_min, workcell_max)
        return out, time.perf_counter_ns() - t0
    if filter_type == "scdf":
        raise NotImplementedError("filter_type='scdf' (collision/filter.hh) is a sequential greedy scan; only 'centervox' runs on the GPU")
    raise ValueError("filter_type must be one of: 'scdf', 'centervox'")
