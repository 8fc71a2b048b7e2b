"""Pins the CPU oracle (oracle/vamp_oracle.c) against the reference's own code compiled in place
(oracle/_ref/libvamp_ref.so).  Skipped where that library has not been built (it needs
/root/reference; the built .so travels to the GPU box)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests import parity, scenes

pytestmark = pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built")

ROBOTS = ["panda", "ur5", "fetch", "baxter"]
KEEP_OUT = {"panda": 0.0, "ur5": 0.0, "fetch": 0.45, "baxter": 0.5}


def both_envs(scene):
    p = scenes.packed(scene)
    return po.add_scene(po.OracleEnv(), p), po.add_scene(po.RefEnv(), p)


@pytest.mark.parametrize("robot", ROBOTS)
def test_sphere_fk(robot):
    o, r = po.Oracle(robot), po.Ref(robot)
    assert (o.dof, o.n_spheres, o.model["resolution"]) == (r.dof, r.n_spheres, r.resolution)
    q = scenes.random_configs(robot, 3000, seed=0)
    fo, fr = o.sphere_fk(q), r.sphere_fk(q)
    # FK sphere centres within 1e-5 m (north_star); we are two orders tighter
    assert np.linalg.norm(fo[..., :3] - fr[..., :3], axis=-1).max() < 2e-6
    assert np.array_equal(fo[..., 3], fr[..., 3])


@pytest.mark.parametrize("seed", range(4))
def test_environment_fields_sort_and_min_distance(seed):
    eo, er = both_envs(scenes.random_scene(seed, n_spheres=7, n_cuboids=8, n_capsules=8))
    for kind in range(5):
        do, dr = eo.dump(kind), er.dump(kind)
        assert do.shape == dr.shape, kind
        if do.size:
            # same classification, same order, min_distance to float rounding
            assert np.array_equal(do[:, :-1], dr[:, :-1]), kind
            assert np.abs(do[:, -1] - dr[:, -1]).max() < 1e-6, kind


@pytest.mark.parametrize("robot", ROBOTS)
def test_config_and_edge_verdicts(robot):
    o, r = po.Oracle(robot), po.Ref(robot)
    named = [scenes.random_scene(s, keep_out=KEEP_OUT[robot]) for s in range(3)] + [{"order": []}]
    if robot == "panda":
        named += [scenes.sphere_cage(), scenes.table_shelf_scene(), scenes.box_scene()]
    for sc in named:
        eo, er = both_envs(sc)
        q = scenes.random_configs(robot, 6000, seed=11)
        vo, vr = o.validate_configs(eo, q), r.validate_configs(er, q, threads=4)
        parity.assert_configs(o, eo, q, vo, vr, "oracle vs reference")
        a, b = scenes.random_edges(robot, 1500, seed=12)
        eo_v, er_v = o.validate_edges(eo, a, b), r.validate_edges(er, a, b, threads=4)
        parity.assert_edges(o, eo, a, b, eo_v, er_v, "oracle vs reference")


def test_sphere_cage_known_answers():
    # scripts/sphere_cage_example.py:10-31,67 of the reference: both endpoints valid, the straight
    # line between them is not
    o = po.Oracle("panda")
    eo, er = both_envs(scenes.sphere_cage())
    q = np.array([scenes.CAGE_A, scenes.CAGE_B], np.float32)
    assert o.validate_configs(eo, q).tolist() == [True, True]
    assert po.Ref("panda").validate_configs(er, q).tolist() == [True, True]
    assert not o.validate_edges(eo, q[:1], q[1:])[0]
    assert not po.Ref("panda").validate_edges(er, q[:1], q[1:])[0]
    fk = o.sphere_fk(q[:1])[0, 58]
    assert np.allclose(fk, [0.3069904, 0.0730000, 0.4878696, 0.012], atol=2e-6)


def test_edge_step_count_matches_reference_control_flow():
    # n = max(ceil(|b-a| / 8 * resolution), 1) with the reference's hsum order (validate.hh:41)
    o = po.Oracle("panda")
    a, b = scenes.random_edges("panda", 200, seed=5)
    for i in range(len(a)):
        # the reference's own arithmetic, step by step in f32: squares, the AVX hsum tree
        # ((l4+l0)+(l6+l2)) + ((l5+l1)+(l7+l3)) (avx.hh:441-452), sqrt, / 8 * resolution, ceil
        v = (b[i] - a[i]).astype(np.float32)
        l = np.zeros(8, np.float32)
        l[:7] = v * v
        s = np.float32(np.float32(np.float32(l[4] + l[0]) + np.float32(l[6] + l[2])) + np.float32(np.float32(l[5] + l[1]) + np.float32(l[7] + l[3])))
        d = np.sqrt(s, dtype=np.float32)
        want = max(int(np.ceil(np.float32(np.float32(d / np.float32(8)) * np.float32(32)))), 1)
        assert o.edge_steps(a[i], b[i]) == want


@pytest.mark.parametrize("robot", ["panda", "fetch"])
def test_capt_pointcloud(robot):
    rng = np.random.default_rng(3)
    o, r = po.Oracle(robot), po.Ref(robot)
    m = o.model
    # points on a few random planes / blobs around the robot
    pts = np.concatenate(
        [
            rng.uniform([0.3, -0.6, 0.0], [0.9, 0.6, 0.02], size=(1500, 3)),
            rng.normal([0.5, 0.3, 0.6], 0.05, size=(700, 3)),
            rng.uniform([-0.8, -0.8, 0.0], [0.8, 0.8, 1.2], size=(300, 3)),
        ]
    ).astype(np.float32)
    if robot == "fetch":
        pts = pts[np.hypot(pts[:, 0], pts[:, 1]) > 0.5]
    eo, er = po.OracleEnv(), po.RefEnv()
    for e in (eo, er):
        e.add_capt(pts, m["min_radius"], m["max_radius"], 0.0025)
    q = scenes.random_configs(robot, 3000, seed=21)
    vo, vr = o.validate_configs(eo, q), r.validate_configs(er, q, threads=4)
    assert 0.02 < vr.mean() < 0.98
    parity.assert_configs(o, eo, q, vo, vr, "capt, oracle vs reference", has_cloud=True)
    a, b = scenes.random_edges(robot, 600, seed=22)
    parity.assert_edges(o, eo, a, b, o.validate_edges(eo, a, b), r.validate_edges(er, a, b, threads=4), "capt, oracle vs reference", has_cloud=True)


@pytest.mark.parametrize("robot", ["panda", "fetch"])
def test_mvt_pointcloud(robot):
    """Multi-level Voxel Table (collision/mvt.hh): the oracle's dense-cell restatement against the
    reference's three-level pointer tables, through the full hierarchy (bounding spheres larger
    than r_max are queried with the +-1 voxel clamp, as the reference does)."""
    rng = np.random.default_rng(5)
    o, r = po.Oracle(robot), po.Ref(robot)
    m = o.model
    pts = np.concatenate(
        [
            rng.uniform([0.3, -0.6, 0.0], [0.9, 0.6, 0.02], size=(1500, 3)),
            rng.normal([0.5, 0.3, 0.6], 0.05, size=(700, 3)),
            rng.uniform([-0.8, -0.8, 0.0], [0.8, 0.8, 1.2], size=(300, 3)),
        ]
    ).astype(np.float32)
    pts = pts[np.hypot(pts[:, 0], pts[:, 1]) > (0.5 if robot == "fetch" else 0.25)]
    lo, hi = [-1.5, -1.5, -0.5], [1.5, 1.5, 2.5]
    pts = pts[np.all((pts > lo) & (pts < hi), axis=1)]
    # the reference's voxels hold at most (r_max / 0.02)^3 points (mvt.hh:455-457, it expects a
    # filtered cloud): keep one point per 2.5 cm cell
    _, keep = np.unique(np.floor(pts / 0.025).astype(np.int64), axis=0, return_index=True)
    pts = pts[np.sort(keep)]
    eo, er = po.OracleEnv(), po.RefEnv()
    for e in (eo, er):
        e.add_mvt(pts, m["min_radius"], m["max_radius"], lo, hi, 0.0025)
    q = scenes.random_configs(robot, 3000, seed=23)
    vo, vr = o.validate_configs(eo, q), r.validate_configs(er, q, threads=4)
    assert 0.02 < vr.mean() < 0.98
    parity.assert_configs(o, eo, q, vo, vr, "mvt, oracle vs reference", has_cloud=True)
    a, b = scenes.random_edges(robot, 600, seed=24)
    parity.assert_edges(o, eo, a, b, o.validate_edges(eo, a, b), r.validate_edges(er, a, b, threads=4), "mvt, oracle vs reference", has_cloud=True)


def test_heightfield():
    rng = np.random.default_rng(4)
    xd, yd = 64, 48
    data = (0.6 * rng.random((yd, xd)) ** 3).astype(np.float32)
    yy, xx = np.mgrid[0:yd, 0:xd]
    data[np.hypot(xx - xd / 2, yy - yd / 2) < 7] = 0.0  # flat ground under the robot base
    f6 = np.array([0.0, 0.0, -0.12, 1 / 0.05, 1 / 0.05, 1 / 1.0], np.float32)  # inverse scales as stored
    o, r = po.Oracle("panda"), po.Ref("panda")
    eo, er = po.OracleEnv(), po.RefEnv()
    for e in (eo, er):
        e.add_heightfield(f6, xd, yd, data.reshape(-1))
    q = scenes.random_configs("panda", 4000, seed=31)
    vo, vr = o.validate_configs(eo, q), r.validate_configs(er, q, threads=4)
    assert 0.02 < vr.mean() < 0.98
    parity.assert_configs(o, eo, q, vo, vr, "heightfield, oracle vs reference")


@pytest.mark.parametrize("robot", ROBOTS)
def test_attachment(robot):
    rng = np.random.default_rng(6)
    o, r = po.Oracle(robot), po.Ref(robot)
    sc = scenes.random_scene(2, keep_out=KEEP_OUT[robot])
    eo, er = both_envs(sc)
    tf12 = np.array([0.0, 0.0, 0.08, 1, 0, 0, 0, 1, 0, 0, 0, 1], np.float32)
    spheres = np.array([[0, 0, 0, 0.04], [0, 0, 0.06, 0.03], [0.03, 0, 0.1, 0.025]], np.float32)
    for e in (eo, er):
        e.attach(tf12, spheres)
    q = scenes.random_configs(robot, 3000, seed=41)
    vo, vr = o.validate_configs(eo, q), r.validate_configs(er, q, threads=1)
    parity.assert_configs(o, eo, q, vo, vr, "attachment, oracle vs reference")
    # the attachment matters: verdicts differ from the detached environment somewhere
    eo.detach()
    assert (o.validate_configs(eo, q) != vo).any()
    a, b = scenes.random_edges(robot, 400, seed=42)
    eo2, er2 = both_envs(sc)
    for e in (eo2, er2):
        e.attach(tf12, spheres)
    parity.assert_edges(o, eo2, a, b, o.validate_edges(eo2, a, b), r.validate_edges(er2, a, b, threads=1), "attachment, oracle vs reference")


@pytest.mark.parametrize("robot", ROBOTS)
def test_eefk_close_to_reference(robot):
    o, r = po.Oracle(robot), po.Ref(robot)
    q = scenes.random_configs(robot, 50, seed=51)
    for i in range(len(q)):
        # the reference's scalar eefk uses libm sin/cos, the traced programs the cephes variant:
        # they differ by the polynomial error (4e-6), not more
        assert np.abs(o.eefk(q[i]) - r.eefk(q[i])).max() < 2e-5


def test_debug_attribution():
    o, r = po.Oracle("panda"), po.Ref("panda")
    eo, er = both_envs(scenes.table_shelf_scene())
    q = scenes.random_configs("panda", 40, seed=61)
    for i in range(len(q)):
        (eh_o, sh_o), (eh_r, sh_r) = o.debug(eo, q[i]), r.debug(er, q[i])
        assert sorted(map(tuple, eh_o)) == sorted(map(tuple, eh_r))
        assert sorted(map(tuple, sh_o)) == sorted(map(tuple, sh_r))


def test_rsqrt_mode_agrees_with_reference_early_out():
    # with the host's RSQRTSS (what the reference executes) the oracle reproduces the reference's
    # early-out decisions exactly; with the exact sqrt it may only test MORE objects
    o, r = po.Oracle("panda"), po.Ref("panda")
    eo, er = both_envs(scenes.table_shelf_scene(seed=3))
    q = scenes.random_configs("panda", 20000, seed=71)
    vr = r.validate_configs(er, q, threads=4)
    o.set_sqrt_mode(1)
    v1 = o.validate_configs(eo, q)
    o.set_sqrt_mode(0)
    v0 = o.validate_configs(eo, q)
    parity.assert_configs(o, eo, q, v1, vr, "rsqrt mode")
    # exact sqrt: every disagreement is "reference skipped an object it should have tested"
    assert not np.any(v0 & ~vr & ~v1)


@pytest.mark.parametrize("robot", ["panda", "fetch"])
def test_filter_self_from_pointcloud(robot):
    """Helper::filter_self_from_pointcloud (bindings/robot_helper.hh:284-322): points overlapping the
    robot at a configuration, or colliding with the environment, are dropped."""
    rng = np.random.default_rng(8)
    o, r = po.Oracle(robot), po.Ref(robot)
    sc = scenes.random_scene(3, keep_out=0.45 if robot == "fetch" else 0.0)
    eo, er = po.add_scene(po.OracleEnv(), scenes.packed(sc)), po.add_scene(po.RefEnv(), scenes.packed(sc))
    pts = rng.uniform([-1.0, -1.0, -0.2], [1.0, 1.0, 1.4], size=(20000, 3)).astype(np.float32)
    for q in scenes.random_configs(robot, 3, seed=12):
        ko, kr = o.filter_points(eo, q, pts, 0.01), r.filter_points(er, q, pts, 0.01)
        assert 0.3 < kr.mean() < 0.999
        bad = np.nonzero(ko != kr)[0]
        if len(bad):
            assert parity.classify(o.point_clearance(eo, q, pts[bad], 0.01), has_cloud=False).all(), bad
