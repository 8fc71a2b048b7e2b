"""Parity of the CUDA path (through the C ABI) with the CPU oracle, the committed golden vectors and
-- where oracle/_ref travelled to the box -- the reference's own compiled code.

Bar (BASELINE.json north_star): verdicts bit-exact except for units whose minimum signed clearance
lies within 1e-5 m of zero (counted and asserted to be the only mismatches); FK sphere centres
within 1e-5 m."""
from pathlib import Path

import numpy as np
import pytest

import vamp_mvt_b200 as vmv
from oracle import pyoracle as po
from tests import parity, scenes
from vamp_mvt_b200 import _lib

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).resolve().parent / "golden"
ROBOTS = ["panda", "ur5", "fetch", "baxter"]
KEEP_OUT = {"panda": 0.0, "ur5": 0.0, "fetch": 0.45, "baxter": 0.5}
BAND = 1e-5  # metres
KINDS = {0: "spheres", 1: "cuboids", 2: "capsules"}


def assert_verdicts(robot, oracle, oenv, q, got, want, what, has_cloud=False):
    """Every mismatching configuration must lie inside the clearance band (tests/parity.py)."""
    return parity.assert_configs(oracle, oenv, q, got, want, what, has_cloud)


def assert_edge_verdicts(oracle, oenv, a, b, got, want, what, has_cloud=False):
    """Every mismatching edge must have a rake state inside the clearance band."""
    return parity.assert_edges(oracle, oenv, a, b, got, want, what, has_cloud)


def scenes_for(robot):
    named = [(f"random{s}", scenes.random_scene(s, keep_out=KEEP_OUT[robot])) for s in (0, 2, 5)] + [("empty", {"order": []})]
    if robot == "panda":
        named += [("cage", scenes.sphere_cage()), ("table", scenes.table_shelf_scene()), ("box", scenes.box_scene())]
    return named


@pytest.mark.parametrize("robot", ROBOTS)
def test_sphere_fk(robot):
    R, O = getattr(vmv, robot), po.Oracle(robot)
    q = scenes.random_configs(robot, 5000, seed=1)
    got, want = R.fk_batch(q), O.sphere_fk(q)
    assert np.linalg.norm(got[..., :3] - want[..., :3], axis=-1).max() < 2e-6  # bar: 1e-5 m
    assert np.array_equal(got[..., 3], want[..., 3])
    single = R.fk(q[0])
    assert len(single) == R.n_spheres() and abs(single[3].x - want[0, 3, 0]) < 2e-6


@pytest.mark.parametrize("path", [1, 2, 3, 0], ids=["per_thread_kernel", "block_kernel", "grid_kernel", "auto_kernel"])
@pytest.mark.parametrize("robot", ROBOTS)
def test_config_verdicts(robot, path):
    R, O = getattr(vmv, robot), po.Oracle(robot)
    ref = po.Ref(robot) if po.ref_available() else None
    _lib.lib().vmv_force_kernel_path(path)
    try:
        for name, sc in scenes_for(robot):
            if path == 3 and not sc["order"]:
                continue  # nothing to rasterise: the grid-culled kernel does not apply
            env = scenes.build_product_env(sc)
            oenv = po.add_scene(po.OracleEnv(), scenes.packed(sc))
            q = scenes.random_configs(robot, 30000, seed=3)
            # a quarter outside the joint limits: the statically pruned self-collision lists do not apply there
            q[:7500] += np.random.default_rng(9).uniform(-0.6, 0.6, size=(7500, q.shape[1])).astype(np.float32)
            got = R.validate_batch(q, env)
            assert_verdicts(robot, O, oenv, q, got, O.validate_configs(oenv, q), f"{name} configs vs oracle")
            if ref is not None:
                renv = po.add_scene(po.RefEnv(), scenes.packed(sc))
                assert_verdicts(robot, O, oenv, q, got, ref.validate_configs(renv, q, threads=8), f"{name} configs vs reference")
    finally:
        _lib.lib().vmv_force_kernel_path(0)


@pytest.mark.parametrize("robot", ["panda", "fetch"])
def test_grid_kernel_wide_masks_and_ragged_sizes(robot):
    """The grid-culled kernels with more than 32 objects (64-bit candidate masks), batch sizes that
    are not multiples of 32, and the fall-back above 64 objects."""
    R, O = getattr(vmv, robot), po.Oracle(robot)
    L = _lib.lib()
    keep = max(KEEP_OUT[robot], 0.3)
    for n_obj, seed in ((48, 11), (64, 12), (80, 13)):
        sc = scenes.random_scene(seed, n_spheres=n_obj // 4, n_cuboids=n_obj // 2, n_capsules=n_obj - n_obj // 4 - n_obj // 2,
                                 lo=(-1.2, -1.2, -0.2), hi=(1.2, 1.2, 1.4), keep_out=keep)
        env = scenes.build_product_env(sc)
        oenv = po.add_scene(po.OracleEnv(), scenes.packed(sc))
        for n in (4097, 20011):
            q = scenes.random_configs(robot, n, seed=seed)
            L.vmv_force_kernel_path(3 if n_obj <= 64 else 0)
            try:
                got = R.validate_batch(q, env)
            finally:
                L.vmv_force_kernel_path(0)
            assert_verdicts(robot, O, oenv, q, got, O.validate_configs(oenv, q), f"{n_obj} objects, n={n}")
        a, b = scenes.random_edges(robot, 2049, seed=seed)
        got = R.validate_motion_batch(a, b, env)
        assert_edge_verdicts(O, oenv, a, b, got, O.validate_edges(oenv, a, b), f"{n_obj} objects, edges")
    if robot == "panda":
        # above 64 objects the grid-culled kernel must refuse when forced
        L.vmv_force_kernel_path(3)
        try:
            with pytest.raises(_lib.VmvError):
                R.validate_batch(scenes.random_configs(robot, 5000, seed=1), env)
        finally:
            L.vmv_force_kernel_path(0)


@pytest.mark.parametrize("path", [1, 2, 3, 0], ids=["per_thread_kernel", "block_kernel", "grid_kernel", "auto_kernel"])
@pytest.mark.parametrize("robot", ROBOTS)
def test_edge_verdicts(robot, path):
    if robot == "baxter" and path == 2:
        pytest.skip("baxter edges have no block kernel")
    try:
        _lib.lib().vmv_force_kernel_path(path)
        _edge_verdicts(robot, path)
    finally:
        _lib.lib().vmv_force_kernel_path(0)


def _edge_verdicts(robot, path=0):
    R, O = getattr(vmv, robot), po.Oracle(robot)
    ref = po.Ref(robot) if po.ref_available() else None
    for name, sc in scenes_for(robot):
        if path == 3 and not sc["order"]:
            continue  # nothing to rasterise: the grid-culled kernel does not apply
        env = scenes.build_product_env(sc)
        oenv = po.add_scene(po.OracleEnv(), scenes.packed(sc))
        a, b = scenes.random_edges(robot, 6000, seed=4)
        got = R.validate_motion_batch(a, b, env)
        assert_edge_verdicts(O, oenv, a, b, got, O.validate_edges(oenv, a, b), f"{name} edges vs oracle")
        if ref is not None:
            renv = po.add_scene(po.RefEnv(), scenes.packed(sc))
            assert_edge_verdicts(O, oenv, a, b, got, ref.validate_edges(renv, a, b, threads=8), f"{name} edges vs reference")


@pytest.mark.parametrize("robot", ROBOTS)
def test_golden_vectors(robot):
    d = np.load(GOLDEN / f"{robot}.npz")
    R, O = getattr(vmv, robot), po.Oracle(robot)
    fk = R.fk_batch(d["q"][:64])
    assert np.linalg.norm(fk[..., :3] - d["fk"][..., :3], axis=-1).max() < 2e-6
    for name in d["scene_names"]:
        env = vmv.Environment()
        packed = {k: d[f"{name}_{k}"] for k in ("spheres", "cuboids", "capsules")}
        L = _lib.lib()
        for kind, i in d[f"{name}_order"]:
            row = np.ascontiguousarray(packed[KINDS[int(kind)]][int(i)], dtype=np.float32)
            fn = {0: L.vmv_env_add_spheres, 1: L.vmv_env_add_cuboids, 2: L.vmv_env_add_capsules}[int(kind)]
            _lib.check(fn(env._h, _lib.ptr(row), 1))
        env._dirty = True
        oenv = po.OracleEnv()
        po.add_scene(oenv, {**{k: list(v) for k, v in packed.items()}, "order": [(KINDS[int(k)], int(i)) for k, i in d[f"{name}_order"]]})
        got = R.validate_batch(d["q"], env)
        assert_verdicts(robot, O, oenv, d["q"], got, d[f"{name}_valid"], f"golden {name}")
        assert (R.validate_motion_batch(d["a"], d["b"], env) != d[f"{name}_edge_valid"]).sum() == 0


def test_sphere_cage_known_answers():
    # reference scripts/sphere_cage_example.py:10-31,67 (BASELINE config 1)
    env = scenes.build_product_env(scenes.sphere_cage())
    assert vmv.panda.validate(scenes.CAGE_A, env) is True
    assert vmv.panda.validate(scenes.CAGE_B, env) is True
    assert vmv.panda.validate_motion(scenes.CAGE_A, scenes.CAGE_B, env) is False
    assert vmv.panda.validate(scenes.CAGE_A) is True  # default: empty environment


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 127, 129, 1000, 4097])
def test_ragged_batch_sizes(n):
    env = scenes.build_product_env(scenes.table_shelf_scene())
    oenv = po.add_scene(po.OracleEnv(), scenes.packed(scenes.table_shelf_scene()))
    O = po.Oracle("panda")
    q = scenes.random_configs("panda", max(n, 1), seed=n)[:n]
    got = vmv.panda.validate_batch(q, env)
    assert got.shape == (n,)
    if n:
        assert np.array_equal(got, O.validate_configs(oenv, q))
        a, b = scenes.random_edges("panda", n, seed=n)
        ge = vmv.panda.validate_motion_batch(a, b, env)
        assert (ge != O.validate_edges(oenv, a, b)).sum() == 0


def test_full_size_properties():
    """BASELINE size (2^20 configs): properties that need no oracle pass over the whole batch."""
    n = 1 << 20
    env = scenes.build_product_env(scenes.table_shelf_scene())
    q = scenes.random_configs("panda", n, seed=0)
    v = vmv.panda.validate_batch(q, env)
    # idempotence
    assert np.array_equal(v, vmv.panda.validate_batch(q, env))
    # a verdict does not depend on the unit's position in the batch or on its block-mates
    perm = np.random.default_rng(0).permutation(n)
    assert np.array_equal(vmv.panda.validate_batch(q[perm], env), v[perm])
    # both kernels agree on every unit
    _lib.lib().vmv_force_kernel_path(1)
    try:
        v1 = vmv.panda.validate_batch(q[: 1 << 18], env)
    finally:
        _lib.lib().vmv_force_kernel_path(0)
    assert np.array_equal(v1, v[: 1 << 18])
    # monotonicity: removing every obstacle can only make configurations valid
    v_empty = vmv.panda.validate_batch(q, vmv.Environment())
    assert not np.any(v & ~v_empty)
    # a zero-length edge is the configuration check (validate.hh:70-77 with vector = 0)
    sub = q[: 1 << 16]
    assert np.array_equal(vmv.panda.validate_motion_batch(sub, sub, env), v[: 1 << 16])
    # oracle spot check on a slice of the big batch
    O = po.Oracle("panda")
    oenv = po.add_scene(po.OracleEnv(), scenes.packed(scenes.table_shelf_scene()))
    sl = slice(500000, 520000)
    assert_verdicts("panda", O, oenv, q[sl], v[sl], O.validate_configs(oenv, q[sl]), "2^20 slice")


def test_edges_full_size_properties():
    n = 1 << 18
    env = scenes.build_product_env(scenes.box_scene())
    a, b = scenes.random_edges("panda", n, seed=0)
    e = vmv.panda.validate_motion_batch(a, b, env)
    assert np.array_equal(e, vmv.panda.validate_motion_batch(a, b, env))
    # the last tine of the first rake block is the goal itself: a valid edge has a valid goal
    vb = vmv.panda.validate_batch(b, env)
    assert not np.any(e & ~vb)
    O = po.Oracle("panda")
    oenv = po.add_scene(po.OracleEnv(), scenes.packed(scenes.box_scene()))
    assert_edge_verdicts(O, oenv, a[:5000], b[:5000], e[:5000], O.validate_edges(oenv, a[:5000], b[:5000]), "2^18 edges, first 5000")


@pytest.mark.parametrize("robot", ["panda", "fetch", "ur5"])
def test_capt_pointcloud_and_heightfield(robot):
    """BASELINE config 4 at test size: CAPT pointcloud + heightfield in one environment."""
    rng = np.random.default_rng(3)
    R, O = getattr(vmv, robot), po.Oracle(robot)
    m = O.model
    pts = np.concatenate(
        [
            rng.uniform([0.3, -0.6, 0.0], [0.9, 0.6, 0.02], size=(6000, 3)),
            rng.normal([0.5, 0.3, 0.6], 0.05, size=(2500, 3)),
            rng.uniform([-0.8, -0.8, 0.0], [0.8, 0.8, 1.2], size=(500, 3)),
        ]
    ).astype(np.float32)
    pts = pts[np.hypot(pts[:, 0], pts[:, 1]) > (0.5 if robot == "fetch" else 0.3)]
    xd, yd = 64, 48
    data = (0.5 * rng.random((yd, xd)) ** 3).astype(np.float32)
    yy, xx = np.mgrid[0:yd, 0:xd]
    data[np.hypot(xx - xd / 2, yy - yd / 2) < 12] = 0.0
    env, oenv = vmv.Environment(), po.OracleEnv()
    env.add_capt_pointcloud(pts, m["min_radius"], m["max_radius"], vmv.POINT_RADIUS)
    oenv.add_capt(pts, m["min_radius"], m["max_radius"], vmv.POINT_RADIUS)
    hf = vmv.make_heightfield([0, 0, -0.25], [0.05, 0.05, 1.0], [xd, yd], data)
    env.add_heightfield(hf)
    oenv.add_heightfield(hf.packed(), xd, yd, data.reshape(-1))
    q = scenes.random_configs(robot, 8000, seed=21)
    got, want = R.validate_batch(q, env), O.validate_configs(oenv, q)
    assert 0.02 < want.mean() < 0.98
    assert_verdicts(robot, O, oenv, q, got, want, "capt+heightfield configs vs oracle", has_cloud=True)
    if po.ref_available():
        renv = po.RefEnv()
        renv.add_capt(pts, m["min_radius"], m["max_radius"], vmv.POINT_RADIUS)
        renv.add_heightfield(hf.packed(), xd, yd, data.reshape(-1))
        assert_verdicts(robot, O, oenv, q, got, po.Ref(robot).validate_configs(renv, q, threads=8), "capt+heightfield configs vs reference", has_cloud=True)
    a, b = scenes.random_edges(robot, 1500, seed=22)
    assert_edge_verdicts(O, oenv, a, b, R.validate_motion_batch(a, b, env), O.validate_edges(oenv, a, b), "capt+heightfield edges", has_cloud=True)


def _capt_cases():
    """Clouds and build parameters that stress what the list-free CAPT query derives from the tree: tiny and
    non-power-of-two clouds, coordinate ties (lattice points, duplicates), leaves that keep their representative
    alone (large r_min), queries wider than the lists reach (r_max below the robot's radii), fat points."""
    rng = np.random.default_rng(11)
    surf = np.concatenate([rng.uniform([0.25, -0.5, 0.0], [0.8, 0.5, 0.03], size=(3000, 3)),
                           rng.normal([0.45, 0.25, 0.55], 0.06, size=(1500, 3)),
                           rng.uniform([-0.7, -0.7, 0.0], [0.7, 0.7, 1.1], size=(300, 3))]).astype(np.float32)
    surf = surf[np.hypot(surf[:, 0], surf[:, 1]) > 0.3]
    lattice = (np.stack(np.meshgrid(np.arange(0.3, 0.75, 0.05), np.arange(-0.4, 0.45, 0.05), np.arange(0.0, 1.0, 0.25), indexing="ij"), -1)
               .reshape(-1, 3).astype(np.float32))
    dup = np.concatenate([surf[:700], surf[:700], surf[:300]])
    yield "two points", np.array([[0.45, 0.1, 0.5], [0.3, -0.3, 0.3]], np.float32), (0.03, 0.1, 0.0025)
    yield "three points", surf[[5, 1900, 3300]], (0.03, 0.1, 0.0025)
    for n in (7, 64, 65, 1000):
        yield f"{n} points", surf[rng.choice(len(surf), n, replace=False)], (0.03, 0.1, 0.0025)
    # (exact coordinate ties are left out: which half a tied point lands in is the sort's choice, std::sort in the
    # reference, and the C oracle's own sort disagrees with it on 8 of 6000 configurations of the plain lattice)
    yield "lattice, jittered by 1e-5", lattice + rng.uniform(-1e-5, 1e-5, size=lattice.shape).astype(np.float32), (0.03, 0.1, 0.0025)
    yield "duplicates", dup, (0.03, 0.1, 0.0025)
    yield "lists shorter than the queries (r_max 0.04)", surf, (0.01, 0.04, 0.0025)
    yield "representative-only leaves (r_min 0.09)", surf, (0.09, 0.1, 0.0025)
    yield "wide lists (r_max 0.3)", surf[::2], (0.01, 0.3, 0.0025)
    yield "fat points (r_point 0.03)", surf, (0.03, 0.1, 0.03)
    yield "zero-radius points", surf, (0.03, 0.1, 0.0)


def test_capt_without_lists_on_adversarial_clouds():
    """The CAPT query evaluates list membership per point instead of storing lists (vmv_device.cuh: capt_member): it must
    reproduce the reference's lists -- including what they MISS -- on clouds and parameters chosen to hit every rule."""
    R, O = vmv.panda, po.Oracle("panda")
    ref = po.Ref("panda") if po.ref_available() else None
    q = scenes.random_configs("panda", 6000, seed=77)
    seen_invalid = 0
    for name, pts, (r_min, r_max, r_point) in _capt_cases():
        env, oenv = vmv.Environment(), po.OracleEnv()
        env.add_capt_pointcloud(pts, r_min, r_max, r_point)
        oenv.add_capt(pts, r_min, r_max, r_point)
        got = R.validate_batch(q, env)
        if ref is not None:
            renv = po.RefEnv()
            renv.add_capt(pts, r_min, r_max, r_point)
            want = ref.validate_configs(renv, q, threads=8)
        else:
            want = O.validate_configs(oenv, q)
        seen_invalid += int((~want).sum())
        # classified on the oracle's restatement of the SAME tree (lists and all): a mismatch must sit on a decision
        # boundary of that structure, not merely near the raw surface
        assert_verdicts("panda", O, oenv, q, got, want, f"list-free CAPT, {name}", has_cloud=True)
        a, b = q[:600], q[600:1200]
        assert_edge_verdicts(O, oenv, a, b, R.validate_motion_batch(a, b, env), O.validate_edges(oenv, a, b), f"list-free CAPT edges, {name}", has_cloud=True)
    assert seen_invalid > 2000
    # a single point (the reference's own query crashes on a tree without a split; the oracle's does not)
    env, oenv = vmv.Environment(), po.OracleEnv()
    one = np.array([[0.45, 0.1, 0.5]], np.float32)
    env.add_capt_pointcloud(one, 0.03, 0.1, 0.0025)
    oenv.add_capt(one, 0.03, 0.1, 0.0025)
    assert_verdicts("panda", O, oenv, q, R.validate_batch(q, env), O.validate_configs(oenv, q), "list-free CAPT, one point", has_cloud=True)
    # two trees and a voxel table in one environment: the nearest-point table spans all of them
    pts = next(c for c in _capt_cases() if c[0] == "1000 points")[1]
    env, oenv = vmv.Environment(), po.OracleEnv()
    for e in (env, oenv):
        add = e.add_capt_pointcloud if e is env else e.add_capt
        add(pts[:500], 0.03, 0.1, 0.0025)
        add(pts[500:] + np.float32(0.01), 0.01, 0.05, 0.01)
    lo, hi = [-1.0, -1.0, -0.2], [1.0, 1.0, 1.4]
    extra = (pts[::3] + np.float32([0.0, 0.02, 0.1])).astype(np.float32)
    env.add_mvt_pointcloud(extra, 0.03, 0.1, lo, hi, 0.0025)
    oenv.add_mvt(extra, 0.03, 0.1, lo, hi, 0.0025)
    assert_verdicts("panda", O, oenv, q, R.validate_batch(q, env), O.validate_configs(oenv, q), "two CAPTs + MVT", has_cloud=True)


@pytest.mark.parametrize("robot", ["fetch", "ur5"])
def test_c4_baseline_size(robot):
    """BASELINE config 4 at its full size: a CAPT over 100 k synthetic surface points plus a 256 x 256 heightfield
    (tests/workloads.py), 2^18 configurations on the GPU; a 2^15 sample against the reference's own CAPT +
    validate compiled in place (oracle/_ref), every mismatch classified by the oracle's clearance accounting on
    the raw cloud; without oracle/_ref a 2^11 sample against the C oracle."""
    from tests import workloads

    R, O = getattr(vmv, robot), po.Oracle(robot)
    env, pts, hf, _ = workloads.c4_environment(robot)
    assert len(pts) > 90_000
    n = 1 << 18
    q = scenes.random_configs(robot, n, seed=0)
    got = R.validate_batch(q, env)
    assert 0.05 < got.mean() < 0.95
    assert np.array_equal(got, R.validate_batch(q, env))  # idempotent
    if po.ref_available():
        ns = 1 << 15
        renv = workloads.c4_checker_env(po.RefEnv, robot, pts, hf)
        want = po.Ref(robot).validate_configs(renv, q[:ns], threads=po.host_threads())
        raw = workloads.c4_checker_env(po.OracleEnv, robot, pts, hf, raw_only=True)
        assert_verdicts(robot, O, raw, q[:ns], got[:ns], want, "C4 full size vs reference", has_cloud=True)
    else:
        ns = 1 << 11
        oenv = workloads.c4_checker_env(po.OracleEnv, robot, pts, hf)
        assert_verdicts(robot, O, oenv, q[:ns], got[:ns], O.validate_configs(oenv, q[:ns]), "C4 full size vs oracle", has_cloud=True)
    # monotone: the heightfield and the cloud can only remove valid configurations
    assert not np.any(got & ~R.validate_batch(q, vmv.Environment()))


@pytest.mark.parametrize("robot", ["panda", "fetch", "ur5"])
def test_mvt_pointcloud(robot):
    """Multi-level Voxel Table pointcloud (reference collision/mvt.hh) next to primitives: configs and
    edges against the oracle and, where it travelled, the compiled reference."""
    rng = np.random.default_rng(6)
    R, O = getattr(vmv, robot), po.Oracle(robot)
    m = O.model
    # a small cloud: the reference sizes its voxel pools for 10 % occupied cells (mvt.hh:455-513) and
    # aborts beyond that
    pts = np.concatenate(
        [
            rng.uniform([0.3, -0.6, 0.0], [0.9, 0.6, 0.02], size=(1500, 3)),
            rng.normal([0.5, 0.3, 0.6], 0.05, size=(700, 3)),
            rng.uniform([-0.8, -0.8, 0.0], [0.8, 0.8, 1.2], size=(300, 3)),
        ]
    ).astype(np.float32)
    pts = pts[np.hypot(pts[:, 0], pts[:, 1]) > (0.5 if robot == "fetch" else 0.3)]
    # one point per 2.5 cm cell: the reference's voxels have a fixed capacity (mvt.hh:455-457)
    _, keep = np.unique(np.floor(pts / 0.025).astype(np.int64), axis=0, return_index=True)
    pts = pts[np.sort(keep)]
    lo, hi = [-1.5, -1.5, -0.5], [1.5, 1.5, 2.5]
    sc = scenes.random_scene(7, n_spheres=2, n_cuboids=2, n_capsules=1, keep_out=max(KEEP_OUT[robot], 0.4))
    env = scenes.build_product_env(sc)
    oenv = po.add_scene(po.OracleEnv(), scenes.packed(sc))
    env.add_mvt_pointcloud(pts, m["min_radius"], m["max_radius"], lo, hi, vmv.POINT_RADIUS)
    oenv.add_mvt(pts, m["min_radius"], m["max_radius"], lo, hi, vmv.POINT_RADIUS)
    q = scenes.random_configs(robot, 8000, seed=25)
    got, want = R.validate_batch(q, env), O.validate_configs(oenv, q)
    assert 0.02 < want.mean() < 0.98
    assert_verdicts(robot, O, oenv, q, got, want, "mvt configs vs oracle", has_cloud=True)
    if po.ref_available():
        renv = po.add_scene(po.RefEnv(), scenes.packed(sc))
        renv.add_mvt(pts, m["min_radius"], m["max_radius"], lo, hi, vmv.POINT_RADIUS)
        assert_verdicts(robot, O, oenv, q, got, po.Ref(robot).validate_configs(renv, q, threads=8), "mvt configs vs reference", has_cloud=True)
    a, b = scenes.random_edges(robot, 1500, seed=26)
    assert_edge_verdicts(O, oenv, a, b, R.validate_motion_batch(a, b, env), O.validate_edges(oenv, a, b), "mvt edges", has_cloud=True)


@pytest.mark.parametrize("robot", ROBOTS)
def test_attachment(robot):
    R, O = getattr(vmv, robot), po.Oracle(robot)
    sc = scenes.random_scene(2, keep_out=KEEP_OUT[robot])
    env = scenes.build_product_env(sc)
    oenv = po.add_scene(po.OracleEnv(), scenes.packed(sc))
    tf = np.eye(4, dtype=np.float32)
    tf[:3, 3] = [0.0, 0.0, 0.08]
    att = vmv.Attachment(tf)
    att.add_spheres([vmv.Sphere([0, 0, 0], 0.04), vmv.Sphere([0, 0, 0.06], 0.03), vmv.Sphere([0.03, 0, 0.1], 0.025)])
    env.attach(att)
    oenv.attach(att.packed_tf12(), att.packed_spheres())
    q = scenes.random_configs(robot, 10000, seed=41)
    got, want = R.validate_batch(q, env), O.validate_configs(oenv, q)
    assert_verdicts(robot, O, oenv, q, got, want, "attached configs vs oracle")
    if po.ref_available():
        renv = po.add_scene(po.RefEnv(), scenes.packed(sc))
        renv.attach(att.packed_tf12(), att.packed_spheres())
        assert_verdicts(robot, O, oenv, q, got, po.Ref(robot).validate_configs(renv, q, threads=8), "attached configs vs reference")
    a, b = scenes.random_edges(robot, 1500, seed=42)
    assert_edge_verdicts(O, oenv, a, b, R.validate_motion_batch(a, b, env), O.validate_edges(oenv, a, b), "attached edges")
    env.detach()
    oenv.detach()
    assert (R.validate_batch(q, env) != got).any()
    assert_verdicts(robot, O, oenv, q, R.validate_batch(q, env), O.validate_configs(oenv, q), "detached")


def test_indexed_edges_and_device_api():
    """PRM-style indexed edge sets through the device-pointer entry points (config 5 format)."""
    L = _lib.lib()
    env = scenes.build_product_env(scenes.box_scene())
    O = po.Oracle("panda")
    oenv = po.add_scene(po.OracleEnv(), scenes.packed(scenes.box_scene()))
    V = scenes.random_configs("panda", 2000, seed=8)
    rng = np.random.default_rng(8)
    pairs = rng.integers(0, len(V), size=(5000, 2)).astype(np.uint32)
    dV, dP = L.vmv_dev_alloc(V.nbytes), L.vmv_dev_alloc(pairs.nbytes)
    dB = L.vmv_dev_alloc((len(pairs) + 31) // 32 * 4)
    _lib.check(L.vmv_memcpy_h2d(dV, _lib.ptr(V), V.nbytes, None))
    _lib.check(L.vmv_memcpy_h2d(dP, _lib.ptr(pairs), pairs.nbytes, None))
    # first use of the environment with this robot also rasterises it (one extra launch, once)
    _lib.check(L.vmv_validate_edges_indexed_dev(vmv.panda.id, env.handle, dV, len(V), dP, len(pairs), 0, dB, None))
    before = L.vmv_launch_count()
    _lib.check(L.vmv_validate_edges_indexed_dev(vmv.panda.id, env.handle, dV, len(V), dP, len(pairs), 0, dB, None))
    assert L.vmv_launch_count() == before + 1
    words = np.zeros((len(pairs) + 31) // 32, np.uint32)
    _lib.check(L.vmv_memcpy_d2h(_lib.ptr(words), dB, words.nbytes, None))
    _lib.check(L.vmv_stream_sync(None))
    got = _lib.unpack_bits(words, len(pairs))
    ea, eb = V[pairs[:, 0]], V[pairs[:, 1]]
    assert_edge_verdicts(O, oenv, ea, eb, got, O.validate_edges(oenv, ea, eb), "indexed edges")
    for p in (dV, dP, dB):
        L.vmv_dev_free(p)


def test_uncommitted_environment_is_an_error():
    L = _lib.lib()
    h = L.vmv_env_create()
    q = np.zeros(7, np.float32)
    w = np.zeros(1, np.uint32)
    assert L.vmv_validate_configs(vmv.panda.id, h, _lib.ptr(q), 1, _lib.ptr(w)) == -3  # VMV_ERR_STATE
    assert b"not committed" in L.vmv_last_error()
    L.vmv_env_destroy(h)


@pytest.mark.parametrize("robot", ROBOTS)
def test_debug_attribution(robot):
    """Robot::fkcc_debug (robots/panda.hh:468-5224 and the other three): per-sphere object names and
    self-colliding sphere pairs, against the oracle and -- where it travelled -- the compiled reference."""
    sc = scenes.table_shelf_scene() if robot == "panda" else scenes.random_scene(4, keep_out=KEEP_OUT[robot])
    env = scenes.build_product_env(sc)
    R, O = getattr(vmv, robot), po.Oracle(robot)
    oenv = po.add_scene(po.OracleEnv(), scenes.packed(sc))
    ref = po.Ref(robot) if po.ref_available() else None
    renv = po.add_scene(po.RefEnv(), scenes.packed(sc)) if ref else None
    names = [shape.name for _, shape in sc["order"]]
    q = scenes.random_configs(robot, 40, seed=61)
    n_env = n_self = 0
    for i in range(len(q)):
        per_sphere, self_pairs = R.debug(q[i], env)
        assert len(per_sphere) == R.n_spheres()
        for checker, cenv in ((O, oenv), (ref, renv)):
            if checker is None:
                continue
            eh, sh = checker.debug(cenv, q[i])
            want = [[] for _ in range(R.n_spheres())]
            for s, obj in eh:
                want[s].append(names[obj])
            assert [sorted(x) for x in per_sphere] == [sorted(x) for x in want], (robot, i)
            assert sorted(self_pairs) == sorted(map(tuple, sh.tolist())), (robot, i)
        n_env += sum(len(x) for x in per_sphere)
        n_self += len(self_pairs)
    assert n_env > 0 and (n_self > 0 or robot == "ur5")


def _product_env_from_packed(scene):
    """Environment from packed shape fields (the golden fixtures store what the reference stores)."""
    L = _lib.lib()
    env = vmv.Environment()
    fn = {"spheres": L.vmv_env_add_spheres, "cuboids": L.vmv_env_add_cuboids, "capsules": L.vmv_env_add_capsules}
    for kind, i in scene["order"]:
        row = np.ascontiguousarray(scene[kind][i], dtype=np.float32)
        _lib.check(fn[kind](env._h, _lib.ptr(row), 1))
    env._dirty = True
    return env


def test_mbm_problems_match_reference():
    """Every start and goal of the 1300 Panda MotionBenchMaker problems the reference ships
    (tests/golden/mbm_panda.npz: reference verdicts, 699 / 700 valid on the classic sets as published in
    the reference's resources/README.md:146), through the C ABI; then the grid-culled kernels on a real
    scene of every problem set with 20 k random configurations and 4 k edges against the oracle."""
    from tests.test_golden import mbm_problems

    O = po.Oracle("panda")
    valid_classic = 0
    seen = set()
    for name, index, scene, start, goal, vs, vg, classic in mbm_problems():
        env = _product_env_from_packed(scene)
        got = vmv.panda.validate_batch(np.stack([start, goal]), env)
        assert (bool(got[0]), bool(got[1])) == (vs, vg), (name, index)
        valid_classic += classic and bool(got[0]) and bool(got[1])
        if name not in seen:
            seen.add(name)
            oenv = po.add_scene(po.OracleEnv(), scene)
            q = scenes.random_configs("panda", 20000, seed=index + 1)
            assert_verdicts("panda", O, oenv, q, vmv.panda.validate_batch(q, env), O.validate_configs(oenv, q), f"MBM {name}")
            a, b = scenes.random_edges("panda", 4000, seed=index + 2)
            assert_edge_verdicts(O, oenv, a, b, vmv.panda.validate_motion_batch(a, b, env), O.validate_edges(oenv, a, b), f"MBM {name} edges")
    assert valid_classic == 699 and len(seen) == 13


@pytest.mark.parametrize("robot", ["panda", "fetch", "baxter"])
def test_filter_self_from_pointcloud(robot):
    """vamp.<robot>.filter_self_from_pointcloud (reference bindings/robot_helper.hh:284-322) against the
    oracle, with primitives and a CAPT pointcloud in the environment."""
    rng = np.random.default_rng(8)
    R, O = getattr(vmv, robot), po.Oracle(robot)
    sc = scenes.random_scene(3, keep_out=KEEP_OUT[robot])
    env = scenes.build_product_env(sc)
    oenv = po.add_scene(po.OracleEnv(), scenes.packed(sc))
    cloud = rng.uniform([0.4, -0.5, 0.0], [0.9, 0.5, 0.3], size=(800, 3)).astype(np.float32)
    m = O.model
    env.add_capt_pointcloud(cloud, m["min_radius"], m["max_radius"], vmv.POINT_RADIUS)
    oenv.add_capt(cloud, m["min_radius"], m["max_radius"], vmv.POINT_RADIUS)
    pts = rng.uniform([-1.0, -1.0, -0.2], [1.0, 1.0, 1.4], size=(50001, 3)).astype(np.float32)
    for q in scenes.random_configs(robot, 2, seed=12):
        want = O.filter_points(oenv, q, pts, 0.01)
        got = R.filter_self_from_pointcloud(pts, 0.01, q, env)
        assert 0.2 < want.mean() < 0.999
        kept = np.zeros(len(pts), bool)
        # the kept points come back in order: recover the mask by a merge walk
        j = 0
        for i in range(len(pts)):
            if j < len(got) and np.array_equal(pts[i], got[j]):
                kept[i] = True
                j += 1
        assert j == len(got)
        bad = np.nonzero(kept != want)[0]
        if len(bad):
            ok = parity.classify(O.point_clearance(oenv, q, pts[bad], 0.01), has_cloud=True)
            assert ok.all(), f"{robot}: filtered points outside the clearance band: {bad[~ok]}"


def test_baseline_sizes_cross_kernel_properties():
    """BASELINE.json sizes (2^20 configurations, 2^18 edges): properties that need no CPU run of the
    full batch -- the three kernel generations agree bit for bit, a zero-length edge is the
    configuration check (validate(q) = validate_motion(q, q), planning/validate.hh:70-77), an edge with
    an invalid end point is invalid -- plus a 2^15 sample against the oracle."""
    L = _lib.lib()
    R, O = vmv.panda, po.Oracle("panda")
    sc = scenes.table_shelf_scene()
    env = scenes.build_product_env(sc)
    oenv = po.add_scene(po.OracleEnv(), scenes.packed(sc))
    q = scenes.random_configs("panda", 1 << 20, seed=77)
    verdicts = {}
    for path in (3, 2):
        L.vmv_force_kernel_path(path)
        try:
            verdicts[path] = R.validate_batch(q, env)
        finally:
            L.vmv_force_kernel_path(0)
    # only states inside the clearance band may differ between kernel generations
    assert_verdicts("panda", O, oenv, q, verdicts[3], verdicts[2], "grid-culled vs block-cooperative kernel, 2^20")
    sub = np.random.default_rng(1).choice(len(q), 1 << 15, replace=False)
    assert_verdicts("panda", O, oenv, q[sub], verdicts[3][sub], O.validate_configs(oenv, q[sub]), "2^20 sample")
    zero = R.validate_motion_batch(q[: 1 << 18], q[: 1 << 18], env)
    assert np.array_equal(zero, verdicts[3][: 1 << 18])
    box = scenes.build_product_env(scenes.box_scene())
    a, b = scenes.random_edges("panda", 1 << 18, seed=78)
    edges = {}
    for path in (3, 2):
        L.vmv_force_kernel_path(path)
        try:
            edges[path] = R.validate_motion_batch(a, b, box)
        finally:
            L.vmv_force_kernel_path(0)
    obox = po.add_scene(po.OracleEnv(), scenes.packed(scenes.box_scene()))
    assert_edge_verdicts(O, obox, a, b, edges[3], edges[2], "grid-culled vs block-cooperative edge kernel, 2^18")
    va, vb = R.validate_batch(a, box), R.validate_batch(b, box)
    # tine 7 of the first rake block is b itself (start + vector * 8/8): an invalid b kills the edge
    assert not (edges[3] & ~vb).any()
    assert 0.2 < edges[3].mean() < 0.9 and va.mean() > 0.2


def test_concurrent_callers_share_an_environment():
    """include/vamp_b200.h: a committed environment may be shared by concurrent calls.  Four host threads
    (ctypes drops the GIL) validate different batches and edge sets against one environment whose voxel
    table is built by whichever call comes first."""
    import threading

    env = scenes.build_product_env(scenes.table_shelf_scene())
    env.commit()
    O = po.Oracle("panda")
    oenv = po.add_scene(po.OracleEnv(), scenes.packed(scenes.table_shelf_scene()))
    results, errors = {}, []

    def work(k):
        try:
            q = scenes.random_configs("panda", 20000 + 64 * k, seed=50 + k)
            a, b = scenes.random_edges("panda", 3000 + 32 * k, seed=60 + k)
            for _ in range(3):
                v, e = vmv.panda.validate_batch(q, env), vmv.panda.validate_motion_batch(a, b, env)
            results[k] = (q, v, a, b, e)
        except Exception as ex:  # surfaced below
            errors.append(ex)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k, (q, v, a, b, e) in results.items():
        assert_verdicts("panda", O, oenv, q, v, O.validate_configs(oenv, q), f"thread {k}")
        assert_edge_verdicts(O, oenv, a, b, e, O.validate_edges(oenv, a, b), f"thread {k} edges")


def test_path_validate_matches_segmentwise_oracle():
    """vamp.<robot>.Path.validate (reference planning/plan.hh:155-168): all consecutive segments valid."""
    O = po.Oracle("panda")
    sc = scenes.box_scene()
    env = scenes.build_product_env(sc)
    oenv = po.add_scene(po.OracleEnv(), scenes.packed(sc))
    rng = np.random.default_rng(3)
    q = scenes.random_configs("panda", 4000, seed=91)
    q = q[O.validate_configs(oenv, q)]
    n_true = 0
    for k in range(40):
        n = int(rng.integers(2, 7))
        start = q[rng.integers(0, len(q))]
        steps = rng.normal(0, 0.08 if k % 2 else 0.4, size=(n - 1, 7)).astype(np.float32)
        pts = np.vstack([start, start + np.cumsum(steps, axis=0)]).astype(np.float32)
        path = vmv.panda.Path(pts)
        want = bool(O.validate_edges(oenv, pts[:-1], pts[1:]).all())
        assert path.validate(env) == want
        n_true += want
        assert path.cost() == pytest.approx(float(np.linalg.norm(pts[1:] - pts[:-1], axis=1).sum()), rel=1e-5)
    assert 0 < n_true < 40
    p2 = vmv.panda.Path([q[0], q[1]])
    p2.subdivide()
    assert len(p2) == 3 and np.allclose(p2[1], 0.5 * (q[0] + q[1]), atol=1e-6)


def test_batched_prm_solves_the_sphere_cage():
    """BASELINE config 1's problem (reference scripts/sphere_cage_example.py) through the batched PRM front-end
    (vmv_prm): the straight line is invalid, the roadmap connects start and goal, and every edge of the returned path
    -- and a sample of the roadmap's edges -- is one the oracle's validate_motion accepts.  (tests/test_planner.py
    pins the same call against the reference's own planner, vertex for vertex.)"""
    env = scenes.build_product_env(scenes.sphere_cage())
    O = po.Oracle("panda")
    oenv = po.add_scene(po.OracleEnv(), scenes.packed(scenes.sphere_cage()))
    res = vmv.panda.prm(scenes.CAGE_A, scenes.CAGE_B, env)
    assert res.path is not None and len(res.path) > 2 and res.iterations > 0
    p = np.stack(res.path)
    assert np.allclose(p[0], scenes.CAGE_A, atol=1e-6) and np.allclose(p[-1], scenes.CAGE_B, atol=1e-6)
    assert O.validate_edges(oenv, p[:-1], p[1:]).all()
    assert vmv.panda.Path(res.path).validate(env)
    rm = vmv.panda.roadmap(scenes.CAGE_A, scenes.CAGE_B, env, max_iterations=3000, max_samples=3000)
    e = rm.edges[np.random.default_rng(0).choice(len(rm.edges), min(2000, len(rm.edges)), replace=False)].astype(np.int64)
    lo, hi = e.min(axis=1), e.max(axis=1)  # the reference validated every edge as (earlier vertex -> later vertex)
    ea, eb = rm.vertices[lo], rm.vertices[hi]
    assert_edge_verdicts(O, oenv, ea, eb, np.ones(len(e), bool), O.validate_edges(oenv, ea, eb), "roadmap edges")


def test_capt_device_build_equals_host_build():
    """The CAPT tables built on the device (segmented bitonic sorts, vmv_capt_build.cuh) against the host build (the reference's
    recursion restated, `VMV_CAPT_HOST_BUILD`): nodes, leaf flags, grid points and grid start tables bit for bit, on clouds
    without coordinate ties (which half a tied point lands in is the sort's choice in both)."""
    import ctypes as C
    import os

    from vamp_mvt_b200 import _lib

    L = _lib.lib()
    rng = np.random.default_rng(5)
    clouds = [rng.random((n, 3)).astype(np.float32) * np.float32([1.5, 1.5, 1.0]) for n in (1, 2, 3, 5, 64, 65, 1000, 2048, 2049, 5000)]
    # 10^5 points with pairwise distinct coordinates on every axis (a uniform float32 sample of this size has hundreds of ties)
    n = 100_000
    clouds.append(np.stack([rng.permutation(n), rng.permutation(n), rng.permutation(n)], -1).astype(np.float32) * np.float32([2e-5, 2e-5, 1e-5]))
    for pts in clouds:
        digests = []
        for host in (False, True):
            if host:
                os.environ["VMV_CAPT_HOST_BUILD"] = "1"
            else:
                os.environ.pop("VMV_CAPT_HOST_BUILD", None)
            try:
                env = vmv.Environment()
                env.add_capt_pointcloud(pts, 0.03, 0.24, vmv.POINT_RADIUS)
                out = (C.c_uint64 * 4)()
                _lib.check(L.vmv_env_capt_digest(env.handle, 0, out))
                digests.append(tuple(out))
            finally:
                os.environ.pop("VMV_CAPT_HOST_BUILD", None)
        assert digests[0] == digests[1], f"{len(pts)} points: device {digests[0]} vs host {digests[1]}"


@pytest.mark.parametrize("robot", ["panda", "fetch"])
def test_attachment_next_to_a_pointcloud(robot):
    """A grasped object in a CAPT + heightfield + primitives environment: configuration batches run the any-environment kernel
    with the attachment as its last phase, edge batches the per-thread kernel; both against the oracle and the reference."""
    rng = np.random.default_rng(9)
    R, O = getattr(vmv, robot), po.Oracle(robot)
    m = O.model
    pts = np.concatenate([rng.uniform([0.3, -0.6, 0.0], [0.9, 0.6, 0.02], size=(4000, 3)),
                          rng.normal([0.55, 0.25, 0.55], 0.06, size=(2000, 3))]).astype(np.float32)
    pts = pts[np.hypot(pts[:, 0], pts[:, 1]) > (0.5 if robot == "fetch" else 0.3)]
    xd, yd = 48, 48
    data = (0.3 * rng.random((yd, xd)) ** 3).astype(np.float32)
    yy, xx = np.mgrid[0:yd, 0:xd]
    data[np.hypot(xx - xd / 2, yy - yd / 2) < 12] = 0.0
    sc = scenes.random_scene(2, keep_out=KEEP_OUT[robot])
    env = scenes.build_product_env(sc)
    oenv = po.add_scene(po.OracleEnv(), scenes.packed(sc))
    renv = po.add_scene(po.RefEnv(), scenes.packed(sc)) if po.ref_available() else None
    hf = vmv.make_heightfield([0, 0, -0.25], [0.05, 0.05, 1.0], [xd, yd], data)
    env.add_capt_pointcloud(pts, m["min_radius"], m["max_radius"], vmv.POINT_RADIUS)
    env.add_heightfield(hf)
    for e in (oenv, renv):
        if e is not None:
            e.add_capt(pts, m["min_radius"], m["max_radius"], vmv.POINT_RADIUS)
            e.add_heightfield(hf.packed(), xd, yd, data.reshape(-1))
    tf = np.eye(4, dtype=np.float32)
    tf[:3, 3] = [0.0, 0.0, 0.08]
    att = vmv.Attachment(tf)
    att.add_spheres([vmv.Sphere([0, 0, 0], 0.04), vmv.Sphere([0, 0, 0.06], 0.03), vmv.Sphere([0.03, 0, 0.1], 0.025)])
    q = scenes.random_configs(robot, 8000, seed=51)
    before = R.validate_batch(q, env)
    assert 0.05 < before.mean() < 0.95
    env.attach(att)
    oenv.attach(att.packed_tf12(), att.packed_spheres())
    got = R.validate_batch(q, env)
    assert (got != before).any() and not np.any(got & ~before)  # the attachment can only remove valid configurations
    assert_verdicts(robot, O, oenv, q, got, O.validate_configs(oenv, q), "attachment + pointcloud vs oracle", has_cloud=True)
    if renv is not None:
        renv.attach(att.packed_tf12(), att.packed_spheres())
        assert_verdicts(robot, O, oenv, q, got, po.Ref(robot).validate_configs(renv, q, threads=8), "attachment + pointcloud vs reference", has_cloud=True)
    assert np.array_equal(got[:512], R.validate_batch(q[:512], env))  # small batches take the per-thread kernel: same verdicts
    a, b = scenes.random_edges(robot, 800, seed=52)
    assert_edge_verdicts(O, oenv, a, b, R.validate_motion_batch(a, b, env), O.validate_edges(oenv, a, b), "attachment + pointcloud edges", has_cloud=True)


def test_capt_tree_equals_the_references():
    """The device-built tree against the REFERENCE's own (its CAPT object read out through oracle/_ref): split values bit for
    bit, on clouds without coordinate ties."""
    import ctypes as C

    from vamp_mvt_b200 import _lib

    if not (po.ref_available() and hasattr(po.ref_lib(), "ref_capt_nlog2")):
        pytest.skip("the compiled reference (oracle/_ref) is not here")
    L = _lib.lib()
    rng = np.random.default_rng(8)
    for n in (2, 3, 17, 300, 1000, 4097, 30000):
        pts = (np.stack([rng.permutation(n), rng.permutation(n), rng.permutation(n)], -1).astype(np.float32) + np.float32(0.25)) * np.float32([3e-5, 2e-5, 1e-5])
        env, renv = vmv.Environment(), po.RefEnv()
        env.add_capt_pointcloud(pts, 0.02, 0.15, vmv.POINT_RADIUS)
        renv.add_capt(pts, 0.02, 0.15, vmv.POINT_RADIUS)
        nlog2, tests = renv.capt_tree()
        k = L.vmv_env_capt_nodes(env.handle, 0, None, 0)
        assert k == 2 * len(tests)
        nodes = np.zeros(k, np.float32)
        L.vmv_env_capt_nodes(env.handle, 0, _lib.ptr(nodes), k)
        ours = nodes[0::2]
        assert np.array_equal(ours.view(np.uint32), tests.view(np.uint32)), f"{n} points: {int((ours.view(np.uint32) != tests.view(np.uint32)).sum())} split values differ"
