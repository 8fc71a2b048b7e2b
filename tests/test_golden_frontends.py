"""Front-end components against the committed fixture tests/golden/frontends.npz (generated from the
reference's own compiled code by tools/make_golden_frontends.py): Halton sampler, Path members, simplify,
CenterVox filter.  The CPU tests need neither a GPU nor the reference (edge checks go to the C oracle, which
tests/test_oracle_vs_reference.py pins); the GPU tests run the product end to end."""
from pathlib import Path

import numpy as np
import pytest

import vamp_mvt_b200 as vmv
from vamp_mvt_b200 import simplify as S
from oracle import pyoracle as po
from tests import scenes
from tests import test_pointcloud_filter as tf
from tests import test_simplify as ts

G = np.load(Path(__file__).resolve().parent / "golden" / "frontends.npz")
ROBOTS = ["panda", "ur5", "fetch", "baxter"]


@pytest.mark.parametrize("robot", ROBOTS)
def test_halton_and_path_members_match_golden(robot):
    R = getattr(vmv, robot)
    h = R.halton()
    assert np.array_equal(h.take(512), G[f"halton_{robot}"])
    assert np.array_equal(h.at(999990 + np.arange(32)), G[f"halton_{robot}_at_999990"])
    for op, arg in ((1, 0), (2, 8), (3, 23)):
        p = R.Path(G[f"path_{robot}_in"])
        {1: p.subdivide, 2: lambda: p.interpolate_to_resolution(arg), 3: lambda: p.interpolate_to_n_states(arg)}[op]()
        assert np.array_equal(p.numpy(), G[f"path_{robot}_op{op}"])
        assert np.float32(p.cost()) == G[f"path_{robot}_op{op}_cost"]


def _simplify_case(c, check=None, env=None):
    robot, scene_name, ops, opts = ts.CASES[c]
    st = ts.make_settings(opts, ops)
    rng = S.StreamRNG(G[f"simplify_{c}_samples"])
    return S.simplify(getattr(vmv, robot), G[f"simplify_{c}_in"], env, st, rng, validate_edges=check)


@pytest.mark.parametrize("c", [int(c) for c in G["simplify_cases"]])
def test_simplify_matches_golden_with_the_oracle_as_edge_checker(c):
    robot, scene_name, _, _ = ts.CASES[c]
    o = po.Oracle(robot)
    oenv = po.add_scene(po.OracleEnv(), scenes.packed(ts.scene_of(scene_name, robot)))
    got = _simplify_case(c, check=lambda a, b: o.validate_edges(oenv, a, b))
    assert np.array_equal(got.path.numpy(), G[f"simplify_{c}_out"])
    assert got.iterations == int(G[f"simplify_{c}_iterations"])


@pytest.mark.parametrize("c", [int(c) for c in G["filter_cases"]])
def test_oracle_filter_matches_golden(c):
    kind, n, vox, rng_, origin, lo, hi = tf.CASES[c]
    p = G[f"filter_{c}_in"]
    assert np.array_equal(p[po.filter_centervox(p, vox, rng_, origin, lo, hi)], G[f"filter_{c}_out"], equal_nan=True)


@pytest.mark.gpu
def test_gpu_frontends_match_golden():
    for c in [int(c) for c in G["simplify_cases"]]:
        robot, scene_name, _, _ = ts.CASES[c]
        got = _simplify_case(c, env=scenes.build_product_env(ts.scene_of(scene_name, robot)))
        assert np.array_equal(got.path.numpy(), G[f"simplify_{c}_out"]), c
    for c in [int(c) for c in G["filter_cases"]]:
        kind, n, vox, rng_, origin, lo, hi = tf.CASES[c]
        got = vmv.filter_pointcloud_centervox(G[f"filter_{c}_in"], vox, rng_, origin, lo, hi)
        assert np.array_equal(got, G[f"filter_{c}_out"], equal_nan=True), c
    for robot in ROBOTS:
        R = getattr(vmv, robot)
        ok, q = R.validate_halton(0, 512, None, return_configs=True)
        assert np.array_equal(q, G[f"halton_{robot}"])
