"""Host-side shape factories mirroring ``vamp.Sphere / Cylinder / Cuboid / make_heightfield /
Attachment`` (reference bindings/environment.cc:24-109,241-269) with the construction math of
collision/factory.hh restated in float32 without Eigen:

  * Euler XYZ -> axes: the reference multiplies ``AngleAxisf(phi,Z) * AngleAxisf(theta,Y) *
    AngleAxisf(rho,X)`` (factory.hh:37-43), which in Eigen is a quaternion product, and rotates the
    unit vectors with it; we do the same quaternion arithmetic in float32.
  * Cylinder(center, euler, r, length): end points = center +- R (0,0,length/2) (factory.hh:160-180),
    xv = p2 - p1, rdv = float(1.0 / |v|^2) (factory.hh:113-124).
  * make_heightfield stores the INVERSE scales (factory.hh:365-386).

These objects only carry the fields the reference's shapes store (collision/shapes.hh); the
min_distance / sort / z-aligned classification happen behind the C ABI at commit time.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

F = np.float32


def _quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return (
        F(aw * bw - ax * bx - ay * by - az * bz),
        F(aw * bx + ax * bw + ay * bz - az * by),
        F(aw * by + ay * bw + az * bx - ax * bz),
        F(aw * bz + az * bw + ax * by - ay * bx),
    )


def _quat_rotate(q, v):
    # Eigen QuaternionBase::_transformVector: v + w*uv + q.vec x uv, uv = 2 q.vec x v
    w, x, y, z = q
    vx, vy, vz = (F(c) for c in v)
    ux = F(F(2) * F(y * vz - z * vy))
    uy = F(F(2) * F(z * vx - x * vz))
    uz = F(F(2) * F(x * vy - y * vx))
    return np.array(
        [
            F(vx + F(w * ux) + F(y * uz - z * uy)),
            F(vy + F(w * uy) + F(z * ux - x * uz)),
            F(vz + F(w * uz) + F(x * uy - y * ux)),
        ],
        dtype=F,
    )


def _quat_matrix(q):
    """Eigen QuaternionBase::toRotationMatrix in float32 (the linear part of ``Translation3f * q``,
    factory.hh:170-174)."""
    w, x, y, z = q
    two = F(2)
    tx, ty, tz = F(two * x), F(two * y), F(two * z)
    twx, twy, twz = F(tx * w), F(ty * w), F(tz * w)
    txx, txy, txz = F(tx * x), F(ty * x), F(tz * x)
    tyy, tyz, tzz = F(ty * y), F(tz * y), F(tz * z)
    one = F(1)
    return np.array(
        [
            [F(one - F(tyy + tzz)), F(txy - twz), F(txz + twy)],
            [F(txy + twz), F(one - F(txx + tzz)), F(tyz - twx)],
            [F(txz - twy), F(tyz + twx), F(one - F(txx + tyy))],
        ],
        dtype=F,
    )


def euler_xyz_quaternion(rho, theta, phi):
    """AngleAxis(phi,Z) * AngleAxis(theta,Y) * AngleAxis(rho,X) as a float32 quaternion (w,x,y,z)."""
    rho, theta, phi = F(rho), F(theta), F(phi)
    h = F(0.5)
    qz = (F(np.cos(F(phi * h))), F(0), F(0), F(np.sin(F(phi * h))))
    qy = (F(np.cos(F(theta * h))), F(0), F(np.sin(F(theta * h))), F(0))
    qx = (F(np.cos(F(rho * h))), F(np.sin(F(rho * h))), F(0), F(0))
    return _quat_mul(_quat_mul(qz, qy), qx)


@dataclass
class Sphere:
    """vamp.Sphere(center, radius) -- bindings/environment.cc:24-42."""

    x: float
    y: float
    z: float
    r: float
    name: str = ""

    def __init__(self, center: Sequence[float], radius: float, name: str = ""):
        self.x, self.y, self.z = (float(F(c)) for c in center)
        self.r = float(F(radius))
        self.name = name

    @property
    def position(self):
        return [self.x, self.y, self.z]

    @property
    def min_distance(self) -> float:
        # shapes.hh:238
        return float(F(np.sqrt(F(F(self.x) ** 2 + F(self.y) ** 2 + F(self.z) ** 2)) - F(self.r)))

    def packed(self) -> np.ndarray:
        return np.array([self.x, self.y, self.z, self.r], dtype=F)


class Cuboid:
    """vamp.Cuboid(center, euler_xyz, half_extents) -- bindings/environment.cc:75-104,
    factory.hh:26-61."""

    def __init__(self, center, euler_xyz, half_extents, name: str = ""):
        q = euler_xyz_quaternion(*euler_xyz)
        a1 = _quat_rotate(q, (1, 0, 0))
        a2 = _quat_rotate(q, (0, 1, 0))
        a3 = _quat_rotate(q, (0, 0, 1))
        self.x, self.y, self.z = (float(F(c)) for c in center)
        (self.axis_1_x, self.axis_1_y, self.axis_1_z) = map(float, a1)
        (self.axis_2_x, self.axis_2_y, self.axis_2_z) = map(float, a2)
        (self.axis_3_x, self.axis_3_y, self.axis_3_z) = map(float, a3)
        self.axis_1_r, self.axis_2_r, self.axis_3_r = (float(F(h)) for h in half_extents)
        self.name = name

    def packed(self) -> np.ndarray:
        return np.array(
            [
                self.x, self.y, self.z,
                self.axis_1_x, self.axis_1_y, self.axis_1_z,
                self.axis_2_x, self.axis_2_y, self.axis_2_z,
                self.axis_3_x, self.axis_3_y, self.axis_3_z,
                self.axis_1_r, self.axis_2_r, self.axis_3_r,
            ],
            dtype=F,
        )


class Cylinder:
    """vamp.Cylinder(center, euler_xyz, radius, length) or Cylinder(endpoint1, endpoint2, radius)
    -- bindings/environment.cc:44-73, factory.hh:101-223.  Used as a capsule by add_capsule."""

    def __init__(self, a, b, c, length=None, name: str = ""):
        if length is None:
            p1 = np.array(a, dtype=F)
            p2 = np.array(b, dtype=F)
            radius = c
        else:
            # factory.hh:166-177: tf = Translation(center) * (AngleAxis products), an isometry whose linear part
            # is the quaternion's rotation MATRIX; the end points are tf * (0, 0, +-length/2)
            R = _quat_matrix(euler_xyz_quaternion(*b))
            centre = np.array(a, dtype=F)
            half = F(F(length) / F(2))
            p1 = np.array([F(F(R[i, 2] * half) + centre[i]) for i in range(3)], dtype=F)
            p2 = np.array([F(F(R[i, 2] * F(-half)) + centre[i]) for i in range(3)], dtype=F)
            radius = c
        v = (p2 - p1).astype(F)
        dot = F(F(v[0] * v[0]) + F(v[1] * v[1]) + F(v[2] * v[2]))
        self.x1, self.y1, self.z1 = map(float, p1)
        self.xv, self.yv, self.zv = map(float, v)
        self.r = float(F(radius))
        with np.errstate(divide="ignore"):
            self.rdv = float(F(1.0 / float(dot))) if dot != 0 else float("inf")
        self.name = name

    @property
    def x2(self):
        return float(F(F(self.x1) + F(self.xv)))

    @property
    def y2(self):
        return float(F(F(self.y1) + F(self.yv)))

    @property
    def z2(self):
        return float(F(F(self.z1) + F(self.zv)))

    def packed(self) -> np.ndarray:
        return np.array([self.x1, self.y1, self.z1, self.xv, self.yv, self.zv, self.r, self.rdv], dtype=F)


class HeightField:
    """Result of make_heightfield (factory.hh:365-423; shapes.hh:244-312)."""

    def __init__(self, center, scaling, dims, data, name: str = ""):
        self.x, self.y, self.z = (float(F(c)) for c in center)
        self.xs, self.ys, self.zs = (float(F(F(1) / F(s))) for s in scaling)
        self.xd, self.yd = int(dims[0]), int(dims[1])
        self.data = np.ascontiguousarray(data, dtype=F).reshape(-1)
        assert self.data.size == self.xd * self.yd
        self.name = name

    def packed(self) -> np.ndarray:
        return np.array([self.x, self.y, self.z, self.xs, self.ys, self.zs], dtype=F)


def make_heightfield(center, scaling, dims, data) -> HeightField:
    return HeightField(center, scaling, dims, data)


class Attachment:
    """vamp.Attachment(tf4x4) -- bindings/environment.cc:241-269, collision/attachments.hh."""

    def __init__(self, tf):
        self.tf = np.array(tf, dtype=F).reshape(4, 4)
        self.spheres: List[Sphere] = []
        self.posed_spheres: List[Sphere] = []

    @property
    def relative_frame(self):
        return self.tf

    def add_sphere(self, sphere: Sphere):
        self.spheres.append(sphere)

    def add_spheres(self, spheres):
        self.spheres.extend(spheres)

    def set_ee_pose(self, tf):
        """Attachment::pose (attachments.hh:43-55): posed = (tf_ee * tf) * sphere."""
        n_tf = (np.array(tf, dtype=F).reshape(4, 4) @ self.tf).astype(F)
        self.posed_spheres = []
        for s in self.spheres:
            p = (n_tf[:3, :3] @ np.array([s.x, s.y, s.z], dtype=F) + n_tf[:3, 3]).astype(F)
            self.posed_spheres.append(Sphere(p, s.r))

    def packed_tf12(self) -> np.ndarray:
        # translation then column-major rotation (vector/math.hh:39-51)
        return np.concatenate([self.tf[:3, 3], self.tf[:3, :3].T.reshape(-1)]).astype(F)

    def packed_spheres(self) -> np.ndarray:
        if not self.spheres:
            return np.zeros((0, 4), dtype=F)
        return np.stack([s.packed() for s in self.spheres]).astype(F)
