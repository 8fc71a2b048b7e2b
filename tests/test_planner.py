"""The batched PRM front-end of the C ABI (vmv_prm, csrc/vmv_planner.cu) against the reference's OWN planner:
tests/golden/planner.npz holds roadmaps and solutions produced by the reference's planning/prm.hh compiled in place
(oracle/_ref, exact brute-force NN stand-in, the reference's Halton sampler; tools/make_planner_golden.py).  The bulk
GPU formulation must return the same vertices, the same adjacency (entry for entry), the same iteration counts and
the same path."""
from pathlib import Path

import numpy as np
import pytest

import vamp_mvt_b200 as vmv
from vamp_mvt_b200 import _lib

GOLDEN = Path(__file__).resolve().parent / "golden" / "planner.npz"
KINDS = {0: "spheres", 1: "cuboids", 2: "capsules"}


def _env(d, key):
    L = _lib.lib()
    env = vmv.Environment()
    fn = {0: L.vmv_env_add_spheres, 1: L.vmv_env_add_cuboids, 2: L.vmv_env_add_capsules}
    for kind, i in d[f"{key}_order"]:
        row = np.ascontiguousarray(d[f"{key}_{KINDS[int(kind)]}"][int(i)], dtype=np.float32)
        _lib.check(fn[int(kind)](env._h, _lib.ptr(row), 1))
    env._dirty = True
    return env


def test_golden_is_committed():
    d = np.load(GOLDEN)
    assert set(d["cases"]) == {"cage", "mbm_bookshelf_small"}
    assert d["cage_vertices"].shape[1] == 7 and len(d["cage_edges"]) > 1000


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["cage", "mbm_bookshelf_small"])
def test_roadmap_is_the_references(key):
    d = np.load(GOLDEN)
    env = _env(d, key)
    rm = vmv.panda.roadmap(d[f"{key}_start"], d[f"{key}_goal"], env, max_iterations=4000, max_samples=4000)
    V, E = d[f"{key}_vertices"], d[f"{key}_edges"]
    assert rm.iterations == int(d[f"{key}_iterations"])
    assert rm.vertices.shape == V.shape and np.array_equal(rm.vertices, V)  # start, goal, the valid Halton samples in order
    # same edge set ...
    want = set(map(tuple, np.sort(E.astype(np.int64), axis=1).tolist()))
    got = set(map(tuple, np.sort(rm.edges.astype(np.int64), axis=1).tolist()))
    assert got == want, (len(got - want), len(want - got))
    # ... and the same adjacency lists, entry for entry (neighbours by distance, then later vertices as they arrive)
    assert np.array_equal(rm.edges, E)
    assert rm.samples_drawn == 4000 and rm.edges_checked >= len(E) // 2


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["cage", "mbm_bookshelf_small"])
def test_prm_solve_is_the_references(key):
    d = np.load(GOLDEN)
    env = _env(d, key)
    res = vmv.panda.prm(d[f"{key}_start"], d[f"{key}_goal"], env)
    P = d[f"{key}_path"]
    assert res.iterations == int(d[f"{key}_solve_iterations"])
    assert res.path is not None and len(res.path) == len(P)
    assert np.array_equal(np.stack(res.path), P)
    assert res.cost == pytest.approx(float(d[f"{key}_cost"]), rel=1e-6)
    assert vmv.panda.Path(res.path).validate(env)


@pytest.mark.gpu
def test_prm_straight_line_and_limits():
    from tests import scenes

    env = vmv.Environment()
    a = np.array(scenes.CAGE_A, np.float32)
    b = a.copy()
    b[0] += 0.4
    assert vmv.panda.validate_motion(a, b, env)
    res = vmv.panda.prm(a, b, env)  # empty environment: the straight line is valid (prm.hh:57-70)
    assert res.iterations == 0 and len(res.path) == 2 and np.array_equal(res.path[0], a) and np.array_equal(res.path[1], b)
    with pytest.raises(_lib.VmvError):
        vmv.panda.roadmap(a, b, env, max_iterations=10**7)  # beyond the exact range of the device Halton sampler


@pytest.mark.gpu
@pytest.mark.parametrize("batch", [1000, 100, 20])
@pytest.mark.parametrize("key", ["cage", "mbm_bookshelf_small"])
def test_fcit_solve_is_the_references(key, batch):
    """FCIT* (planning/fcit.hh): host search over GPU sample batches and GPU edge rows -- the reference's path, cost
    and iteration count."""
    d = np.load(GOLDEN)
    env = _env(d, key)
    tag = "fcit" if batch == 1000 else f"fcit{batch}"
    res = vmv.panda.fcit(d[f"{key}_start"], d[f"{key}_goal"], env, batch_size=batch, max_samples=100000 if batch == 1000 else 4000)
    P = d[f"{key}_{tag}_path"]
    assert res.iterations == int(d[f"{key}_{tag}_iterations"])
    assert res.path is not None and len(res.path) == len(P) and np.array_equal(np.stack(res.path), P)
    assert res.cost == pytest.approx(float(d[f"{key}_{tag}_cost"]), rel=1e-6)
    assert vmv.panda.Path(res.path).validate(env)
    assert res.edges_checked > 0
