"""Oracle against the committed golden vectors (tests/golden/*.npz, generated from the reference's own
compiled code by tools/make_golden.py).  Runs anywhere -- no GPU, no reference needed."""
from pathlib import Path

import numpy as np
import pytest

from oracle import pyoracle as po

GOLDEN = Path(__file__).resolve().parent / "golden"
ROBOTS = ["panda", "ur5", "fetch", "baxter"]
KINDS = {0: "spheres", 1: "cuboids", 2: "capsules"}


def load_scene(d, name):
    p = {k: [row for row in d[f"{name}_{k}"]] for k in ("spheres", "cuboids", "capsules")}
    p["order"] = [(KINDS[int(k)], int(i)) for k, i in d[f"{name}_order"]]
    return p


@pytest.mark.parametrize("robot", ROBOTS)
def test_oracle_matches_golden(robot):
    d = np.load(GOLDEN / f"{robot}.npz")
    o = po.Oracle(robot)
    fk = o.sphere_fk(d["q"][:64])
    assert np.linalg.norm(fk[..., :3] - d["fk"][..., :3], axis=-1).max() < 2e-6
    assert np.array_equal(fk[..., 3], d["fk"][..., 3])
    for name in d["scene_names"]:
        env = po.add_scene(po.OracleEnv(), load_scene(d, name))
        v = o.validate_configs(env, d["q"])
        bad = np.nonzero(v != d[f"{name}_valid"])[0]
        if len(bad):
            assert np.abs(o.min_clearance(env, d["q"][bad])).max() <= 1e-5, (robot, name)
        e = o.validate_edges(env, d["a"], d["b"])
        assert (e != d[f"{name}_edge_valid"]).sum() == 0, (robot, name)


def test_sphere_cage_golden():
    from tests import scenes

    d = np.load(GOLDEN / "sphere_cage.npz")
    assert d["validate"].tolist() == [True, True] and d["validate_motion"].tolist() == [False]
    o = po.Oracle("panda")
    env = po.add_scene(po.OracleEnv(), scenes.packed(scenes.sphere_cage()))
    q = np.stack([d["a"], d["b"]])
    assert o.validate_configs(env, q).tolist() == [True, True]
    assert o.validate_edges(env, q[:1], q[1:]).tolist() == [False]
    assert np.abs(o.sphere_fk(q[:1])[0] - d["fk_a"]).max() < 2e-6


# ---- MotionBenchMaker problems shipped with the reference (resources/panda/problems.tar.bz2) ----------
def mbm_problems():
    """Yields (set name, index, packed scene dict, start, goal, ref valid_start, ref valid_goal, classic)
    from tests/golden/mbm_panda.npz (tools/make_mbm_golden.py: the reference's converter rules and the
    reference's own verdicts)."""
    z = np.load(GOLDEN / "mbm_panda.npz")
    d = {k: z[k] for k in z.files}  # (an NpzFile decompresses a member on every access)
    kinds = {0: "spheres", 1: "cuboids", 2: "capsules"}
    for k in range(len(d["start"])):
        lo, hi = d["offsets"][k], d["offsets"][k + 1]
        scene = {"spheres": [], "cuboids": [], "capsules": [], "order": []}
        for kind, i in d["order"][lo:hi]:
            name = kinds[int(kind)]
            scene["order"].append((name, len(scene[name])))
            scene[name].append(d[name][int(i)])
        yield (str(d["sets"][d["set_id"][k]]), int(d["index"][k]), scene, d["start"][k], d["goal"][k],
               bool(d["valid_start"][k]), bool(d["valid_goal"][k]), bool(d["classic"][k]))


def test_mbm_problems_oracle_matches_reference_and_published_count():
    """All 1300 Panda MBM problems: the oracle's verdict for every start and goal equals the
    reference's, and the seven classic sets give the published 699 valid of 700
    (reference resources/README.md:146)."""
    oracle = po.Oracle("panda")
    n = valid_classic = n_classic = 0
    for name, index, scene, start, goal, vs, vg, classic in mbm_problems():
        env = po.add_scene(po.OracleEnv(), scene)
        got = oracle.validate_configs(env, np.stack([start, goal]))
        assert (bool(got[0]), bool(got[1])) == (vs, vg), (name, index)
        n += 1
        n_classic += classic
        valid_classic += classic and vs and vg
    assert n == 1300 and n_classic == 700 and valid_classic == 699
