"""``Environment`` mirroring ``vamp.Environment`` (reference bindings/environment.cc:111-181).

Shapes are forwarded to the C ABI as they are added; the device copy is (re)built lazily at the
first validation call after a change (``commit``)."""
from __future__ import annotations

import ctypes as C
import time
from typing import Optional

import numpy as np

from . import _lib
from .shapes import Attachment, Cuboid, Cylinder, HeightField, Sphere


class Environment:
    def __init__(self):
        self._L = _lib.lib()
        self._h = C.c_void_p(self._L.vmv_env_create())
        if not self._h:
            raise _lib.VmvError("vmv_env_create failed")
        self._dirty = True
        self.names = []  # object id (insertion order) -> name, used by debug()
        self.attachment: Optional[Attachment] = None

    def __del__(self):
        try:
            self._L.vmv_env_destroy(self._h)
        except Exception:
            pass

    # -- adders, same names and meaning as the reference -----------------------------------
    def add_sphere(self, sphere: Sphere):
        a = sphere.packed()
        _lib.check(self._L.vmv_env_add_spheres(self._h, _lib.ptr(a), 1))
        self.names.append(sphere.name)
        self._dirty = True

    def add_cuboid(self, cuboid: Cuboid):
        a = cuboid.packed()
        _lib.check(self._L.vmv_env_add_cuboids(self._h, _lib.ptr(a), 1))
        self.names.append(cuboid.name)
        self._dirty = True

    def add_capsule(self, cylinder: Cylinder):
        a = cylinder.packed()
        _lib.check(self._L.vmv_env_add_capsules(self._h, _lib.ptr(a), 1))
        self.names.append(cylinder.name)
        self._dirty = True

    def add_heightfield(self, hf: HeightField):
        a = hf.packed()
        _lib.check(self._L.vmv_env_add_heightfield(self._h, _lib.ptr(a), hf.xd, hf.yd, _lib.ptr(hf.data)))
        self.names.append(hf.name)
        self._dirty = True

    def add_capt_pointcloud(self, points, r_min: float, r_max: float, r_point: float) -> int:
        """Returns the build time in nanoseconds like the reference (environment.cc:150-160)."""
        p = _lib.f32(points).reshape(-1, 3)
        t0 = time.perf_counter_ns()
        _lib.check(self._L.vmv_env_add_capt(self._h, _lib.ptr(p), len(p), r_min, r_max, r_point))
        self.names.append("")
        self._dirty = True
        return time.perf_counter_ns() - t0

    def add_mvt_pointcloud(self, points, r_min: float, r_max: float, aabb_min, aabb_max, r_point: float) -> int:
        """Multi-level Voxel Table pointcloud (environment.cc:163-176); returns the build time in ns."""
        p = _lib.f32(points).reshape(-1, 3)
        lo, hi = _lib.f32(aabb_min).reshape(3), _lib.f32(aabb_max).reshape(3)
        t0 = time.perf_counter_ns()
        _lib.check(self._L.vmv_env_add_mvt(self._h, _lib.ptr(p), len(p), r_min, r_max, _lib.ptr(lo), _lib.ptr(hi), r_point))
        self.names.append("")
        self._dirty = True
        return time.perf_counter_ns() - t0

    # upstream VAMP calls this add_pointcloud; the fork splits it into capt / mvt
    add_pointcloud = add_capt_pointcloud

    def attach(self, attachment: Attachment):
        s = attachment.packed_spheres()
        tf = attachment.packed_tf12()
        _lib.check(self._L.vmv_env_attach(self._h, _lib.ptr(tf), _lib.ptr(s), len(s)))
        self.attachment = attachment
        self._dirty = True

    def detach(self):
        _lib.check(self._L.vmv_env_detach(self._h))
        self.attachment = None
        self._dirty = True

    # -- engine side ------------------------------------------------------------------------
    def commit(self):
        if self._dirty:
            _lib.check(self._L.vmv_env_commit(self._h))
            self._dirty = False
        return self

    @property
    def handle(self):
        self.commit()
        return self._h

    def dump(self, kind: int) -> np.ndarray:
        width = {0: 5, 1: 9, 2: 9, 3: 16, 4: 16}[kind]
        n = _lib.check(self._L.vmv_env_dump(self._h, kind, None, 0))
        buf = np.zeros(max(1, n * width), np.float32)
        self._L.vmv_env_dump(self._h, kind, _lib.ptr(buf), buf.size)
        return buf[: n * width].reshape(n, width)
