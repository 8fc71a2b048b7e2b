"""Batched path simplification (vamp_mvt_b200/simplify.py) against the reference's own
``simplify<Robot, 8, resolution>`` (planning/simplify.hh:191-258) compiled in place (oracle/_ref).

CPU tests check the host logic: the batched routines are given the compiled reference's
``validate_motion`` as edge checker, so any difference from the reference's result is a difference in the
restated control flow / arithmetic / random draws.  The GPU tests run the product end to end."""
import numpy as np
import pytest

import vamp_mvt_b200 as vmv
from vamp_mvt_b200 import simplify as S
from oracle import pyoracle as po
from tests import scenes

pytestmark = pytest.mark.skipif(not po.ref_available() or not hasattr(po.ref_lib(), "ref_simplify"),
                                reason="oracle/_ref not built (or built without simplify)")


def jagged_path(robot: str, ref, renv, seed: int, n_waypoints: int = 14) -> np.ndarray:
    """A valid but wasteful path: a random walk whose every segment the reference accepts."""
    rng = np.random.default_rng(seed)
    qs = scenes.random_configs(robot, 4000, seed=seed)
    qs = qs[ref.validate_configs(renv, qs)]
    path = [qs[0]]
    m = scenes.robot_model(robot)
    lo = np.array(m["lower"], np.float32)
    hi = lo + np.array(m["range"], np.float32)
    while len(path) < n_waypoints:
        d = rng.normal(size=(64, m["dof"]))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        cand = np.clip(path[-1] + d * rng.uniform(0.4, 1.6, size=(64, 1)), lo, hi).astype(np.float32)
        ok = ref.validate_edges(renv, np.repeat(path[-1][None, :], 64, 0), cand)
        if ok.any():
            path.append(cand[np.argmax(ok)])
    return np.stack(path)


def settings12(st: S.SimplifySettings):
    return [st.max_iterations, st.interpolate, st.bspline.max_steps, st.bspline.min_change,
            st.bspline.midpoint_interpolation, st.reduce.max_steps, st.reduce.max_empty_steps, st.reduce.range_ratio,
            st.perturb.max_steps, st.perturb.max_empty_steps, st.perturb.perturbation_attempts, st.perturb.range]


CASES = [
    ("panda", "cage", [S.SHORTCUT, S.BSPLINE], {}),
    ("panda", "table", [S.SHORTCUT, S.BSPLINE], {}),
    ("panda", "box", [S.BSPLINE], {"bspline_steps": 3}),
    ("panda", "table", [S.SHORTCUT], {}),
    ("ur5", "table", [S.SHORTCUT, S.BSPLINE], {"interpolate": 40}),
    ("fetch", "random", [S.SHORTCUT, S.BSPLINE], {}),
    ("panda", "table", [S.REDUCE], {}),
    ("panda", "box", [S.REDUCE, S.SHORTCUT, S.BSPLINE], {"interpolate": 30}),
    ("panda", "table", [S.PERTURB], {}),
    ("panda", "box", [S.BSPLINE, S.PERTURB, S.SHORTCUT, S.REDUCE], {"bspline_steps": 2, "midpoint": 0.3}),
    ("baxter", "random", [S.SHORTCUT, S.BSPLINE, S.PERTURB], {}),
]


def scene_of(name, robot):
    if name == "cage":
        return scenes.sphere_cage()
    if name == "table":
        return scenes.table_shelf_scene()
    if name == "box":
        return scenes.box_scene()
    return scenes.random_scene(5, keep_out={"fetch": 0.45, "baxter": 0.8}.get(robot, 0.0))


def make_settings(opts, ops):
    st = S.SimplifySettings(operations=list(ops))
    st.interpolate = opts.get("interpolate", 0)
    st.bspline.max_steps = opts.get("bspline_steps", 1)
    st.bspline.midpoint_interpolation = opts.get("midpoint", 0.5)
    return st


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_batched_simplify_reproduces_the_reference_path(case, seed):
    robot, scene_name, ops, opts = CASES[case]
    ref = po.Ref(robot)
    renv = po.add_scene(po.RefEnv(), scenes.packed(scene_of(scene_name, robot)))
    path = jagged_path(robot, ref, renv, seed)
    st = make_settings(opts, ops)
    samples = np.random.default_rng(100 + seed).random((257, ref.dof), dtype=np.float32)

    want, want_it = ref.simplify(renv, path, ops, settings12(st), samples)

    calls = []

    def check(a, b):
        calls.append(len(a))
        return ref.validate_edges(renv, a, b)

    got = S.simplify(getattr(vmv, robot), path, None, st, S.StreamRNG(samples), validate_edges=check)
    g = got.path.numpy()
    assert g.shape == want.shape, (g.shape, want.shape)
    assert np.array_equal(g, want)  # bit-exact waypoints
    assert got.iterations == want_it
    assert len(want) >= 2 and np.array_equal(want[0], path[0]) and np.array_equal(want[-1], path[-1])
    if S.SHORTCUT in ops and len(calls) > 1:
        assert max(calls) > 1  # candidates really went out as batches


def test_distribution_matches_libstdcxx_stream():
    # REDUCE alone draws only from Distribution; a long run through the reference pins the integer stream
    robot = "panda"
    ref = po.Ref(robot)
    renv = po.add_scene(po.RefEnv(), scenes.packed(scenes.table_shelf_scene()))
    path = jagged_path(robot, ref, renv, 3, n_waypoints=40)
    st = S.SimplifySettings(operations=[S.REDUCE], max_iterations=8)
    st.reduce.max_steps, st.reduce.max_empty_steps = 60, 30
    want, _ = ref.simplify(renv, path, [S.REDUCE], settings12(st))
    got = S.simplify(vmv.panda, path, None, st, S.StreamRNG(), validate_edges=lambda a, b: ref.validate_edges(renv, a, b))
    assert np.array_equal(got.path.numpy(), want)


@pytest.mark.parametrize("robot", ["panda", "ur5", "fetch", "baxter"])
def test_halton_sampler_matches_reference(robot):
    ref = po.Ref(robot)
    h = getattr(vmv, robot).halton()
    want = ref.halton(5000)
    assert np.array_equal(h.take(3000), want[:3000])
    assert np.array_equal(np.stack([h.next() for _ in range(50)]), want[3000:3050])
    assert np.array_equal(h.peek(7), want[3050:3057])
    # across the reset / base rotation after 10^6 samples (halton.hh:78-85) and the second one
    for skip in (999990, 2000000 - 20):
        h.reset()
        h.advance(skip)
        assert np.array_equal(h.take(64), ref.halton(64, skip=skip))


def test_perturb_with_the_halton_sampler():
    # the reference's perturb_path scales the (already scaled) Halton sample once more (simplify.hh:171-172)
    ref = po.Ref("panda")
    renv = po.add_scene(po.RefEnv(), scenes.packed(scenes.box_scene()))
    path = jagged_path("panda", ref, renv, 5)
    ops = [S.PERTURB, S.SHORTCUT]
    st = make_settings({}, ops)
    # drive the compiled reference with the unit-cube stream that makes its StreamRNG return Halton's values
    lo, rg = np.array(vmv.panda.lower_bounds(), np.float32), vmv.panda._range
    hal = ref.halton(400)
    want, _ = ref.simplify(renv, path, ops, settings12(st), hal)
    got = S.simplify(vmv.panda, path, None, st, vmv.panda.halton(), validate_edges=lambda a, b: ref.validate_edges(renv, a, b))
    assert np.array_equal(got.path.numpy(), want)


@pytest.mark.parametrize("robot", ["panda", "ur5", "fetch", "baxter"])
def test_path_members_match_reference(robot):
    # Path::cost / subdivide / interpolate_to_resolution / interpolate_to_n_states (planning/plan.hh:12-153)
    ref = po.Ref(robot)
    if not hasattr(ref.lib, "ref_path_op"):
        pytest.skip("oracle/_ref built without path_op")
    R = getattr(vmv, robot)
    for seed in range(4):
        wp = scenes.random_configs(robot, 3 + 2 * seed, seed=40 + seed)
        for op, arg in ((0, 0), (1, 0), (2, 4), (2, 32), (3, 25), (3, 3), (3, 200)):
            want, want_cost = ref.path_op(op, wp, arg)
            p = R.Path(wp)
            if op == 1:
                p.subdivide()
            elif op == 2:
                p.interpolate_to_resolution(arg)
            elif op == 3:
                p.interpolate_to_n_states(arg)
            assert np.array_equal(p.numpy(), want), (robot, seed, op, arg)
            assert np.float32(p.cost()) == np.float32(want_cost)
    p = R.Path()
    p.append(list(wp[0]))
    p.insert(0, wp[1])
    assert len(p) == 2 and p[0].dtype == np.float32 and np.array_equal(p[0], wp[1])


def test_trivial_paths():
    ref = po.Ref("panda")
    renv = po.add_scene(po.RefEnv(), scenes.packed(scenes.sphere_cage()))
    check = lambda a, b: ref.validate_edges(renv, a, b)
    a, b = np.array(scenes.CAGE_A, np.float32), np.array(scenes.CAGE_B, np.float32)
    two = S.simplify(vmv.panda, [a, b], None, None, None, validate_edges=check)
    assert len(two.path) == 2 and two.iterations == 0
    # straight line valid -> endpoints only (simplify.hh:217-225)
    c = a.copy()
    c[0] += np.float32(0.3)
    assert check(a[None, :], c[None, :])[0]
    r = S.simplify(vmv.panda, [a, S.interpolate(a, c, 0.5), c], None, None, None, validate_edges=check)
    assert len(r.path) == 2 and np.array_equal(r.path[0], a) and np.array_equal(r.path[1], c)
    for p in ([a], []):
        assert len(S.simplify(vmv.panda, p, None, None, None, validate_edges=check).path) == len(p)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [0, 1, 2, 4, 7, 9])
def test_gpu_simplify_matches_reference(case):
    robot, scene_name, ops, opts = CASES[case]
    ref = po.Ref(robot)
    scene = scene_of(scene_name, robot)
    renv = po.add_scene(po.RefEnv(), scenes.packed(scene))
    env = scenes.build_product_env(scene)
    R = getattr(vmv, robot)
    exact = 0
    for seed in range(3):
        path = jagged_path(robot, ref, renv, seed)
        st = make_settings(opts, ops)
        samples = np.random.default_rng(100 + seed).random((257, ref.dof), dtype=np.float32)
        want, _ = ref.simplify(renv, path, ops, settings12(st), samples)
        got = R.simplify(path, env, st, S.StreamRNG(samples))
        g = got.path.numpy()
        # the result must be a valid path of the reference whatever happens inside the clearance band
        assert ref.validate_edges(renv, g[:-1], g[1:]).all()
        assert np.array_equal(g[0], path[0]) and np.array_equal(g[-1], path[-1])
        exact += int(g.shape == want.shape and np.array_equal(g, want))
    # identical verdicts (outside the 1e-5 m band) give the identical path
    assert exact == 3


@pytest.mark.gpu
def test_gpu_plan_then_simplify_on_the_sphere_cage():
    """BASELINE config 1's script (reference scripts/sphere_cage_example.py: plan, then simplify) on the
    batched front-ends: PRM over Halton samples, then simplify; the simplified path must be the one the
    reference's simplify makes of the same planned path."""
    ref = po.Ref("panda")
    scene = scenes.sphere_cage()
    renv = po.add_scene(po.RefEnv(), scenes.packed(scene))
    env = scenes.build_product_env(scene)
    rm = vmv.panda.prm(scenes.CAGE_A, scenes.CAGE_B, env, max_samples=30000)
    assert rm.path is not None and len(rm.path) > 2
    planned = np.stack(rm.path)
    assert ref.validate_edges(renv, planned[:-1], planned[1:]).all()
    st = S.SimplifySettings()
    got = vmv.panda.simplify(planned, env, st, vmv.panda.halton())
    want, _ = ref.simplify(renv, planned, st.operations, settings12(st))
    g = got.path.numpy()
    assert np.array_equal(g, want)
    assert got.cost <= float(np.linalg.norm(planned[1:] - planned[:-1], axis=1).sum()) + 1e-4


def test_device_halton_limits_and_errors():
    # no GPU needed: argument checks come first, and the exact range is a host-side table
    assert vmv.panda.halton_exact_limit() == 1000000 and vmv.fetch.halton_exact_limit() == 1000000
    assert vmv.baxter.halton_exact_limit() == 29 ** 4 - 1
    L = vmv._lib.lib()
    assert L.vmv_halton_fill_dev(vmv.panda.id, 999999, 2, None, None) != 0
    assert L.vmv_halton_fill_dev(vmv.baxter.id, 29 ** 4 - 10, 64, 16, None) != 0
    assert b"exact range" in L.vmv_last_error()


@pytest.mark.gpu
@pytest.mark.parametrize("robot", ["panda", "ur5", "fetch", "baxter"])
def test_gpu_halton_sampler_and_fused_validation(robot):
    """vmv_validate_halton: the configurations the device generates are the reference's Halton samples bit
    for bit, and the verdicts are those of the same configurations uploaded from the host."""
    ref = po.Ref(robot)
    R = getattr(vmv, robot)
    scene = scene_of("table" if robot in ("panda", "ur5") else "random", robot)
    env = scenes.build_product_env(scene)
    limit = R.halton_exact_limit()
    for first, n in ((0, 5000), (123457, 4096), (limit - 3000, 3000), (7, 1), (0, 33)):
        ok, q = R.validate_halton(first, n, env, return_configs=True)
        assert np.array_equal(q, ref.halton(n, skip=first))          # the compiled reference's sampler
        assert np.array_equal(q, R.halton().at(first + np.arange(n)))  # the host restatement
        assert np.array_equal(ok, R.validate_batch(q, env))
    with pytest.raises(Exception):
        R.validate_halton(limit - 10, 11, env)


@pytest.mark.gpu
def test_gpu_prm_vertices_are_the_valid_samples_of_the_host_halton_stream():
    """The planner samples on the device; its vertices must be the valid members of the host restatement's stream, in order."""
    env = scenes.build_product_env(scenes.sphere_cage())
    a = vmv.panda.prm(scenes.CAGE_A, scenes.CAGE_B, env, max_samples=30000)
    assert a.path is not None and a.samples_drawn > 0
    q = vmv.panda.halton().at(np.arange(a.samples_drawn))
    ok = vmv.panda.validate_batch(q, env)
    want = q[ok]
    got = a.vertices[2:]
    assert len(got) <= len(want) and np.array_equal(got, want[: len(got)])
    assert np.array_equal(a.vertices[0], np.asarray(scenes.CAGE_A, np.float32)) and np.array_equal(a.vertices[1], np.asarray(scenes.CAGE_B, np.float32))
