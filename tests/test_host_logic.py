"""CPU-side tests: the C-ABI library loads and exports every declared symbol, shape factories,
environment packing (min_distance / classification / sort) against the oracle, error behaviour
without a GPU, robot constants and the robot model files."""
import ctypes
import json
import re
from pathlib import Path

import numpy as np
import pytest

import vamp_mvt_b200 as vmv
from oracle import pyoracle as po
from tests import scenes
from vamp_mvt_b200 import _lib, sharding

REPO = Path(__file__).resolve().parents[1]


def test_abi_exports_every_declared_symbol():
    header = (REPO / "include" / "vamp_b200.h").read_text()
    names = set(re.findall(r"\b(vmv_[a-z0-9_]+)\s*\(", header))
    assert len(names) > 30
    L = ctypes.CDLL(str(_lib.LIB_PATH))
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing


def test_library_is_cuda_native():
    # the product library must not link the oracle or any CPU fallback
    import subprocess

    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "or_validate" not in out and "ref_validate" not in out
    sass = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass


def test_robot_constants_match_reference_headers():
    # reference robots/panda.hh:14-18,50-66 etc.
    expect = {"panda": (7, 59, 32), "ur5": (6, 40, 32), "fetch": (8, 111, 32), "baxter": (14, 75, 64)}
    for name, (dof, ns, res) in expect.items():
        r = getattr(vmv, name)
        assert (r.dimension(), r.n_spheres(), r.resolution()) == (dof, ns, res)
        assert len(r.joint_names()) == dof
    assert np.allclose(vmv.panda.lower_bounds(), [-2.9671, -1.8326, -2.9671, -3.1416, -2.9671, -0.0873, -2.9671])
    assert np.allclose(vmv.panda.upper_bounds(), [2.9671, 1.8326, 2.9671, 0.0873, 2.9671, 3.8223, 2.9671], atol=1e-6)
    assert vmv.panda.min_max_radii() == pytest.approx((0.012, 0.08), abs=1e-6)
    assert vmv.panda.end_effector() == "panda_grasptarget"


def test_shape_factories():
    # z rotations keep the cuboid z-aligned exactly (axis_3_z == 1), as in the reference's quaternion path
    c = vmv.Cuboid([0.5, 0.1, 0.3], [0, 0, 0.7], [0.1, 0.2, 0.3])
    assert c.axis_3_z == 1.0 and c.axis_3_x == 0.0
    assert c.axis_1_x == pytest.approx(np.cos(0.7), abs=1e-6) and c.axis_1_y == pytest.approx(np.sin(0.7), abs=1e-6)
    # general rotation: R = Rz(phi) Ry(theta) Rx(rho), axes are its columns (factory.hh:37-43)
    rho, theta, phi = 0.3, -0.5, 1.1
    Rx = np.array([[1, 0, 0], [0, np.cos(rho), -np.sin(rho)], [0, np.sin(rho), np.cos(rho)]])
    Ry = np.array([[np.cos(theta), 0, np.sin(theta)], [0, 1, 0], [-np.sin(theta), 0, np.cos(theta)]])
    Rz = np.array([[np.cos(phi), -np.sin(phi), 0], [np.sin(phi), np.cos(phi), 0], [0, 0, 1]])
    R = Rz @ Ry @ Rx
    c = vmv.Cuboid([0, 0, 0], [rho, theta, phi], [1, 1, 1])
    got = np.array([[c.axis_1_x, c.axis_2_x, c.axis_3_x], [c.axis_1_y, c.axis_2_y, c.axis_3_y], [c.axis_1_z, c.axis_2_z, c.axis_3_z]])
    assert np.abs(got - R).max() < 1e-6
    # capsule from centre/euler/length: endpoints centre +- R (0,0,len/2), rdv = 1/|v|^2 (factory.hh:113-124,160-180)
    k = vmv.Cylinder([0.1, 0.2, 0.3], [rho, theta, phi], 0.05, 0.4)
    p1 = np.array([0.1, 0.2, 0.3]) + R @ [0, 0, 0.2]
    assert np.abs(np.array([k.x1, k.y1, k.z1]) - p1).max() < 1e-6
    assert np.abs(np.array([k.xv, k.yv, k.zv]) - R @ [0, 0, -0.4]).max() < 1e-6
    assert k.rdv == pytest.approx(1 / 0.16, rel=1e-5)
    # upright capsule is z-aligned exactly
    k = vmv.Cylinder([0.1, 0.2, 0.3], [0, 0, 0], 0.05, 0.4)
    assert k.xv == 0.0 and k.yv == 0.0
    k2 = vmv.Cylinder([0, 0, 0], [0, 0, 1], 0.1)
    assert (k2.zv, k2.rdv) == (1.0, 1.0)
    assert vmv.Sphere([3, 4, 0], 1).min_distance == pytest.approx(4.0)
    hf = vmv.make_heightfield([0, 0, 0], [0.5, 0.25, 2.0], [4, 3], np.zeros(12))
    assert (hf.xs, hf.ys, hf.zs) == (2.0, 4.0, 0.5)


def test_shape_factories_match_the_reference_factory_golden():
    """tests/golden/factory.npz (tools/make_factory_golden.py): the fields the reference's own factory.hh stores
    for Cuboid(center, euler, half), Cylinder(center, euler, r, length), Cylinder(p1, p2, r) and make_heightfield --
    factory.hh compiled in place (oracle/ref/ref_factory.cc) -- plus euler_matrix(rho, theta, phi, 'sxyz') of the
    reference's vendored transformations.py.  Conventions must agree exactly (which axis is which column, end point
    order, inverse scales, the z-aligned predicates bindings/environment.cc:124,138 read); values to a few ulp
    (libm vs numpy sin/cos, and the FMA contraction the reference's flags leave to the compiler)."""
    d = np.load(REPO / "tests" / "golden" / "factory.npz")
    ulp1 = float(np.spacing(np.float32(1)))
    for i in range(len(d["centre"])):
        c, e = d["centre"][i], d["euler"][i]
        box = vmv.Cuboid(c, e, d["half"][i]).packed()
        want = d["cuboid"][i]
        assert np.array_equal(box[:3], want[:3]) and np.array_equal(box[12:], want[12:])
        assert np.abs(box[3:12] - want[3:12]).max() <= 4 * ulp1, i
        assert (box[11] == 1.0) == (want[11] == 1.0)  # z-aligned classification of the cuboid
        axes = np.stack([box[3:6], box[6:9], box[9:12]], axis=1)  # columns
        assert np.abs(axes - d["euler_matrix"][i]).max() < 1e-6, i
        cyl, want = vmv.Cylinder(c, e, d["radius"][i], d["length"][i]).packed(), d["cylinder_center"][i]
        assert np.abs(cyl[:6] - want[:6]).max() <= 8 * ulp1 * max(1.0, float(np.abs(want[:6]).max())), i
        assert cyl[6] == want[6]
        assert ((cyl[3] == 0) and (cyl[4] == 0)) == ((want[3] == 0) and (want[4] == 0))  # z-aligned capsule
        # rdv is float(1.0 / |v|^2) of the stored v on both sides (factory.hh:113-124)
        for k in (cyl, want):
            v = k[3:6].astype(np.float64)
            assert k[7] == pytest.approx(1.0 / float(v @ v), rel=1e-6)
        # end point order: p1 = centre + R (0, 0, +length/2)
        assert np.abs((cyl[:3] + 0.5 * cyl[3:6]) - c).max() < 1e-6
        assert float(cyl[3:6] @ d["euler_matrix"][i][:, 2]) == pytest.approx(-float(d["length"][i]), rel=1e-5)
        cyl, want = vmv.Cylinder(c, d["p2"][i], d["radius"][i]).packed(), d["cylinder_endpoints"][i]
        assert np.array_equal(cyl[:7], want[:7]) and cyl[7] == pytest.approx(want[7], rel=4 * ulp1), i
        hf = vmv.make_heightfield(c, d["scale"][i], [2, 2], np.zeros(4)).packed()
        assert np.array_equal(hf, d["heightfield"][i])


@pytest.mark.parametrize("seed", range(3))
def test_environment_packing_matches_oracle(seed):
    sc = scenes.random_scene(seed, n_spheres=6, n_cuboids=9, n_capsules=7)
    env = scenes.build_product_env(sc)
    oenv = po.add_scene(po.OracleEnv(), scenes.packed(sc))
    total = 0
    for kind in range(5):
        a, b = env.dump(kind), oenv.dump(kind)
        assert a.shape == b.shape
        total += len(a)
        if a.size:
            assert np.array_equal(a[:, :-1], b[:, :-1])
            assert np.abs(a[:, -1] - b[:, -1]).max() < 1e-6
            assert np.all(np.diff(a[:, -1]) >= 0)  # ascending min_distance
    assert total == len(sc["order"])


def test_problem_dict_to_vamp_rules():
    # src/vamp/__init__.py:141-188 of the reference: cylinders -> capsules, except "box" -> cuboids
    prob = {
        "problem": "table_pick",
        "sphere": [{"name": "s", "position": [1, 0, 0], "radius": 0.1}],
        "cylinder": [{"name": "c", "position": [0.5, 0, 0.5], "orientation_euler_xyz": [0, 0, 0], "radius": 0.05, "length": 0.3}],
        "box": [{"name": "b", "position": [0.5, 0.5, 0.5], "orientation_euler_xyz": [0, 0, 0.2], "half_extents": [0.1, 0.1, 0.1]}],
    }
    env = vmv.problem_dict_to_vamp(prob)
    assert [len(env.dump(k)) for k in range(5)] == [1, 0, 1, 0, 1]
    env = vmv.problem_dict_to_vamp(dict(prob, problem="box"))
    assert [len(env.dump(k)) for k in range(5)] == [1, 0, 0, 0, 2]
    env = vmv.problem_dict_to_vamp(prob, ignore_names=["b", "s"])
    assert [len(env.dump(k)) for k in range(5)] == [0, 0, 1, 0, 0]


def test_argument_errors_and_no_cpu_fallback():
    L = _lib.lib()
    assert L.vmv_robot_id(b"nonexistent") < 0
    assert b"unknown robot" in L.vmv_last_error()
    with pytest.raises(ValueError):
        vmv.panda.validate([0.0] * 6)
    env = vmv.Environment()
    if L.vmv_device_count() == 0:
        # no GPU here: every compute entry point must fail loudly, never fall back to the CPU
        with pytest.raises(_lib.VmvError):
            vmv.panda.validate([0.0] * 7, env)
        with pytest.raises(_lib.VmvError):
            vmv.panda.fk([0.0] * 7)
        with pytest.raises(_lib.VmvError):
            vmv.panda.validate_motion_batch(np.zeros((4, 7)), np.ones((4, 7)), env)


def test_bounds_check_semantics():
    # Helper::validate (robot_helper.hh:255-267): out of bounds + check_bounds -> False without touching the GPU
    q = np.array(vmv.panda.upper_bounds()) + 0.1
    assert vmv.panda.in_bounds(vmv.panda.lower_bounds())
    assert not vmv.panda.in_bounds(q)
    assert vmv.panda.validate(q, vmv.Environment(), check_bounds=True) is False


def test_robot_models_are_consistent():
    for name in vmv.ROBOT_NAMES:
        m = json.loads((REPO / "vamp_mvt_b200" / "robots" / f"{name}.json").read_text())
        L = m["links"]
        assert sum(len(l["spheres"]) for l in L) == m["n_spheres"]
        for l in L:
            # our bounding sphere encloses every fine sphere of the link
            c, R = np.array(l["bound"][:3]), l["bound"][3]
            for s in l["spheres"]:
                assert np.linalg.norm(np.array(s[:3]) - c) + s[3] <= R + 1e-9
        for e in m["self_pair_info"]:
            na, nb = len(L[e["a"]]["spheres"]), len(L[e["b"]]["spheres"])
            assert L[e["a"]]["body"] != L[e["b"]]["body"]
            if e["pruned"] is not None:
                assert all(0 <= i < na and 0 <= j < nb for i, j in e["pruned"])


def test_eefk_host():
    T = vmv.panda.eefk(scenes.CAGE_A)
    assert T.shape == (4, 4) and np.allclose(T[3], [0, 0, 0, 1])
    assert np.allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-5)
    o = po.Oracle("panda")
    assert np.abs(T - o.eefk(scenes.CAGE_A)).max() < 2e-5


def test_shard_bounds_cover_and_align():
    for n in [0, 1, 31, 32, 33, 1000, 1 << 20, 10**8]:
        for world in [1, 2, 4, 8]:
            spans = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (lo, hi), (lo2, _) in zip(spans, spans[1:]):
                assert hi == lo2 and lo % 32 == 0 and lo2 % 32 == 0 or hi == n


@pytest.mark.gpu
def test_capt_host_build_takes_any_cloud():
    """`add_capt_pointcloud` is host work (the k-d tree of capt.hh:106-119, no affordance lists): it must accept the clouds the
    reference accepts -- empty, tiny, non-power-of-two, duplicated, with coordinate ties -- in milliseconds, and refuse what
    its 24-bit leaf tags cannot address.  (Marked gpu because an environment belongs to a device from its first shape on.)"""
    import time

    rng = np.random.default_rng(0)
    clouds = [np.zeros((0, 3), np.float32), rng.random((1, 3)), rng.random((2, 3)), rng.random((3, 3)), rng.random((1000, 3)),
              np.repeat(rng.random((10, 3)), 7, axis=0), np.round(rng.random((500, 3)), 1)]
    for pts in clouds:
        env = vmv.Environment()
        env.add_capt_pointcloud(pts.astype(np.float32), 0.01, 0.08, 0.0025)
        assert vmv.panda.validate_batch(np.zeros((64, 7), np.float32), env).shape == (64,)  # commits and queries the tree
    big = (rng.random((100_000, 3)) * [2.0, 2.0, 1.0]).astype(np.float32)
    env = vmv.Environment()
    t0 = time.perf_counter()
    env.add_capt_pointcloud(big, 0.03, 0.24, 0.0025)
    assert time.perf_counter() - t0 < 5.0  # 0.05 s on 16 threads; the lists of the reference take 9 s at this r_max
    L = _lib.lib()
    few = np.zeros((4, 3), np.float32)
    rc = L.vmv_env_add_capt(env.handle, _lib.ptr(few), (1 << 24) + 1, 0.01, 0.08, 0.0025)  # refused before the points are read
    assert rc < 0 and b"2^24" in L.vmv_last_error()
