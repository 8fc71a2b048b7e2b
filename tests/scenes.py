"""Seeded scenes and inputs shared by the tests, the smoke check and bench.py.

A scene is a dict {'order': [(kind, shape), ...]} of host-side shape objects from
vamp_mvt_b200.shapes, in insertion order.  ``packed`` turns it into the plain field arrays the CPU
checkers take, ``build_product_env`` into a vamp_mvt_b200.Environment.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from vamp_mvt_b200.shapes import Cuboid, Cylinder, Sphere

REPO = Path(__file__).resolve().parents[1]

# scripts/sphere_cage_example.py:10-31 of the reference (BASELINE config 1)
CAGE_A = [0.0, -0.785, 0.0, -2.356, 0.0, 1.571, 0.785]
CAGE_B = [2.35, 1.0, 0.0, -0.8, 0, 2.5, 0.785]
CAGE_CENTRES = [
    [0.55, 0, 0.25], [0.35, 0.35, 0.25], [0, 0.55, 0.25], [-0.55, 0, 0.25], [-0.35, -0.35, 0.25],
    [0, -0.55, 0.25], [0.35, -0.35, 0.25], [0.35, 0.35, 0.8], [0, 0.55, 0.8], [-0.35, 0.35, 0.8],
    [-0.55, 0, 0.8], [-0.35, -0.35, 0.8], [0, -0.55, 0.8], [0.35, -0.35, 0.8],
]


def sphere_cage(radius: float = 0.2):
    return {"order": [("sphere", Sphere(c, radius, name=f"cage{i}")) for i, c in enumerate(CAGE_CENTRES)]}


def table_shelf_scene(seed: int = 0, n_boxes: int = 12, n_cylinders: int = 4):
    """Synthetic MotionBenchMaker-style table + shelf scene of cuboids and cylinders (BASELINE
    config 2): a table top with legs in front of the robot, a shelf unit with boards and side
    walls, and a few objects (boxes, upright and tilted cylinders) standing on them."""
    rng = np.random.default_rng(seed)
    order = []
    # table: top + 4 legs (axis-aligned boxes, yaw-rotated as a group)
    yaw = rng.uniform(-0.3, 0.3)
    c, s = np.cos(yaw), np.sin(yaw)

    def place(local, centre):
        return [centre[0] + c * local[0] - s * local[1], centre[1] + s * local[0] + c * local[1], centre[2] + local[2]]

    tc = [0.75, 0.0, 0.0]
    order.append(("cuboid", Cuboid(place([0, 0, 0.38], tc), [0, 0, yaw], [0.35, 0.6, 0.02], name="table_top")))
    for i, (lx, ly) in enumerate([(-0.3, -0.55), (-0.3, 0.55), (0.3, -0.55), (0.3, 0.55)]):
        order.append(("cuboid", Cuboid(place([lx, ly, 0.18], tc), [0, 0, yaw], [0.02, 0.02, 0.18], name=f"table_leg{i}")))
    # shelf to the left: back, two sides, three boards
    sy = rng.uniform(0.1, 0.3)
    sc = [0.1, 0.85, 0.0]
    c, s = np.cos(sy + np.pi / 2), np.sin(sy + np.pi / 2)
    order.append(("cuboid", Cuboid(place([0.2, 0, 0.6], sc), [0, 0, sy + np.pi / 2], [0.01, 0.4, 0.6], name="shelf_back")))
    order.append(("cuboid", Cuboid(place([0, -0.4, 0.6], sc), [0, 0, sy + np.pi / 2], [0.2, 0.01, 0.6], name="shelf_side0")))
    order.append(("cuboid", Cuboid(place([0, 0.4, 0.6], sc), [0, 0, sy + np.pi / 2], [0.2, 0.01, 0.6], name="shelf_side1")))
    for i, h in enumerate([0.3, 0.7, 1.1]):
        order.append(("cuboid", Cuboid(place([0, 0, h], sc), [0, 0, sy + np.pi / 2], [0.2, 0.4, 0.01], name=f"shelf_board{i}")))
    # loose boxes (arbitrary orientation) and cylinders on the table
    n_loose = max(0, n_boxes - 11)
    for i in range(n_loose + 1):
        pos = [rng.uniform(0.5, 1.0), rng.uniform(-0.5, 0.5), 0.4 + rng.uniform(0.05, 0.12)]
        order.append(("cuboid", Cuboid(pos, rng.uniform(-0.4, 0.4, 3), rng.uniform(0.03, 0.08, 3), name=f"box{i}")))
    for i in range(n_cylinders):
        pos = [rng.uniform(0.5, 1.0), rng.uniform(-0.5, 0.5), 0.4 + 0.1]
        euler = [0, 0, 0] if i % 2 == 0 else list(rng.uniform(-0.5, 0.5, 3))
        order.append(("capsule", Cylinder(pos, euler, rng.uniform(0.02, 0.05), 0.2, name=f"cyl{i}")))
    return {"order": order}


def box_scene(seed: int = 0):
    """Synthetic MBM-'box'-style scene (BASELINE config 3): an open box (bottom + 4 walls) on a
    pedestal in front of the robot and one cylinder-as-cuboid inside
    (src/vamp/__init__.py:153-171 turns the 'box' problem's cylinders into cuboids)."""
    rng = np.random.default_rng(seed)
    yaw = rng.uniform(-0.4, 0.4)
    c, s = np.cos(yaw), np.sin(yaw)
    centre = [0.6, rng.uniform(-0.15, 0.15), 0.25]

    def place(local):
        return [centre[0] + c * local[0] - s * local[1], centre[1] + s * local[0] + c * local[1], centre[2] + local[2]]

    order = [
        ("cuboid", Cuboid(place([0, 0, 0]), [0, 0, yaw], [0.2, 0.3, 0.01], name="box_bottom")),
        ("cuboid", Cuboid(place([0.2, 0, 0.1]), [0, 0, yaw], [0.01, 0.3, 0.1], name="box_wall0")),
        ("cuboid", Cuboid(place([-0.2, 0, 0.1]), [0, 0, yaw], [0.01, 0.3, 0.1], name="box_wall1")),
        ("cuboid", Cuboid(place([0, 0.3, 0.1]), [0, 0, yaw], [0.2, 0.01, 0.1], name="box_wall2")),
        ("cuboid", Cuboid(place([0, -0.3, 0.1]), [0, 0, yaw], [0.2, 0.01, 0.1], name="box_wall3")),
        ("cuboid", Cuboid([centre[0], centre[1], 0.12], [0, 0, yaw], [0.25, 0.35, 0.12], name="pedestal")),
        ("cuboid", Cuboid(place([0.05, 0.05, 0.08]), [0, 0, 0], [0.03, 0.03, 0.07], name="can")),
    ]
    return {"order": order}


def random_scene(seed: int, n_spheres=5, n_cuboids=6, n_capsules=4, lo=(-1, -1, 0.0), hi=(1, 1, 1.2), keep_out=0.0):
    """Mixed primitives, half of the cuboids/capsules z-aligned; objects whose centre is closer than
    keep_out to the z axis are pushed outwards (robots with a wide base)."""
    rng = np.random.default_rng(seed)

    def centre():
        p = rng.uniform(lo, hi)
        d = np.hypot(p[0], p[1])
        if d < keep_out:
            p[:2] *= keep_out / max(d, 1e-3)
        return p

    order = []
    for i in range(n_spheres):
        order.append(("sphere", Sphere(centre(), rng.uniform(0.04, 0.18), name=f"s{i}")))
    for i in range(n_cuboids):
        e = rng.uniform(-np.pi, np.pi, 3) if i % 2 else [0, 0, rng.uniform(-np.pi, np.pi)]
        order.append(("cuboid", Cuboid(centre(), e, rng.uniform(0.03, 0.22, 3), name=f"c{i}")))
    for i in range(n_capsules):
        e = rng.uniform(-np.pi, np.pi, 3) if i % 2 else [0, 0, 0]
        order.append(("capsule", Cylinder(centre(), e, rng.uniform(0.03, 0.09), rng.uniform(0.1, 0.6), name=f"k{i}")))
    perm = rng.permutation(len(order))
    return {"order": [order[i] for i in perm]}


def packed(scene):
    """-> dict for oracle.pyoracle.add_scene."""
    out = {"spheres": [], "cuboids": [], "capsules": [], "order": []}
    key = {"sphere": "spheres", "cuboid": "cuboids", "capsule": "capsules"}
    for kind, shape in scene["order"]:
        k = key[kind]
        out["order"].append((k, len(out[k])))
        out[k].append(shape.packed())
    return out


def build_product_env(scene):
    import vamp_mvt_b200 as vmv

    env = vmv.Environment()
    for kind, shape in scene["order"]:
        {"sphere": env.add_sphere, "cuboid": env.add_cuboid, "capsule": env.add_capsule}[kind](shape)
    return env


def robot_model(robot: str) -> dict:
    return json.loads((REPO / "vamp_mvt_b200" / "robots" / f"{robot}.json").read_text())


def random_configs(robot: str, n: int, seed: int = 0) -> np.ndarray:
    """q_j = s_a[j] + s_m[j] * u, u ~ U[0,1)."""
    m = robot_model(robot)
    rng = np.random.default_rng(seed)
    lo = np.array(m["lower"], np.float32)
    rg = np.array(m["range"], np.float32)
    return (lo + rg * rng.random((n, m["dof"]), dtype=np.float32)).astype(np.float32)


def random_edges(robot: str, n: int, seed: int = 0, lmin: float = 0.25, lmax: float = 2.0):
    """a as random_configs, b = a + L u, u uniform on the unit sphere, L ~ U(lmin, lmax), clipped
    to the joint bounds."""
    m = robot_model(robot)
    rng = np.random.default_rng(seed + 7919)
    a = random_configs(robot, n, seed)
    d = rng.normal(size=(n, m["dof"]))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    L = rng.uniform(lmin, lmax, size=(n, 1))
    lo = np.array(m["lower"], np.float32)
    hi = lo + np.array(m["range"], np.float32)
    b = np.clip(a + d * L, lo, hi).astype(np.float32)
    return a, b
