"""Generates tests/golden/factory.npz: outputs of the reference's own shape factories
(src/impl/vamp/collision/factory.hh compiled in place into oracle/_ref, see oracle/ref/ref_factory.cc) and of the
reference's vendored src/vamp/transformations.py (euler_matrix, 'sxyz' = Rz(phi) Ry(theta) Rx(rho), the
convention factory.hh:37-43 spells with AngleAxis products) on seeded inputs.  Run where /root/reference is
mounted; tests/test_host_logic.py compares vamp_mvt_b200/shapes.py with the committed file."""
import ctypes as C
import importlib.util
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from oracle import pyoracle as po  # noqa: E402


def main():
    po.build()
    L = po.ref_lib()
    spec = importlib.util.spec_from_file_location("ref_transformations", "/root/reference/src/vamp/transformations.py")
    tfm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tfm)

    rng = np.random.default_rng(0)
    n = 200
    centre = rng.uniform(-1.5, 1.5, size=(n, 3)).astype(np.float32)
    euler = rng.uniform(-np.pi, np.pi, size=(n, 3)).astype(np.float32)
    # special orientations: identity, pure yaw (z-aligned shapes), quarter turns, gimbal lock
    special = np.array(
        [[0, 0, 0], [0, 0, 0.7], [0, 0, -2.1], [0, 0, np.pi / 2], [np.pi / 2, 0, 0], [0, np.pi / 2, 0], [np.pi, 0, 0], [0.3, np.pi / 2, -0.4],
         [1e-4, 0, 0.5], [0, 1e-4, 0.5], [np.pi, np.pi, np.pi]], np.float32)
    euler[: len(special)] = special
    half = rng.uniform(0.01, 0.6, size=(n, 3)).astype(np.float32)
    radius = rng.uniform(0.01, 0.3, size=n).astype(np.float32)
    length = rng.uniform(0.02, 1.5, size=n).astype(np.float32)
    p2 = (centre + rng.normal(0, 0.4, size=(n, 3))).astype(np.float32)
    p2[:5, :2] = centre[:5, :2]  # z-aligned capsules from end points
    scale = rng.uniform(0.01, 2.0, size=(n, 3)).astype(np.float32)

    fp = lambda a: a.ctypes.data_as(C.c_void_p)
    cub = np.zeros((n, 15), np.float32)
    cyl_c = np.zeros((n, 8), np.float32)
    cyl_e = np.zeros((n, 8), np.float32)
    hf = np.zeros((n, 6), np.float32)
    rot = np.zeros((n, 3, 3), np.float64)
    for i in range(n):
        L.ref_factory_cuboid(fp(centre[i]), fp(euler[i]), fp(half[i]), fp(cub[i]))
        L.ref_factory_cylinder_center(fp(centre[i]), fp(euler[i]), C.c_float(radius[i]), C.c_float(length[i]), fp(cyl_c[i]))
        L.ref_factory_cylinder_endpoints(fp(centre[i]), fp(p2[i]), C.c_float(radius[i]), fp(cyl_e[i]))
        L.ref_factory_heightfield(fp(centre[i]), fp(scale[i]), fp(hf[i]))
        rot[i] = tfm.euler_matrix(float(euler[i, 0]), float(euler[i, 1]), float(euler[i, 2]), "sxyz")[:3, :3]
    out = REPO / "tests" / "golden" / "factory.npz"
    np.savez_compressed(out, centre=centre, euler=euler, half=half, radius=radius, length=length, p2=p2, scale=scale,
                        cuboid=cub, cylinder_center=cyl_c, cylinder_endpoints=cyl_e, heightfield=hf, euler_matrix=rot)
    print("wrote", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
