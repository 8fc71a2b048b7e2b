"""CPU pin of the host-side FCIT* restatement (vamp_mvt_b200/csrc/vmv_fcit.hpp): driven by the reference's own
fkcc / validate_motion / Halton sampler as providers (oracle/ref/ref_robot.hh, planner mode 3) it must return what the
reference's own FCIT<Robot, 8, resolution>::solve (mode 2) returns -- path, cost, iteration count -- on the sphere
cage and on MotionBenchMaker scenes, across batch sizes that force anything from 2 to 200+ search iterations.  The
GPU suite (tests/test_planner.py) then checks the same search over the GPU providers against committed goldens."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests import scenes

pytestmark = pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built")


_CASES = []


def _cases(limit=5):
    from tests.test_golden import mbm_problems

    if not _CASES:
        _CASES.append(("cage", scenes.packed(scenes.sphere_cage()), scenes.CAGE_A, scenes.CAGE_B))
        seen = set()
        for name, index, scene, start, goal, vs, vg, classic in mbm_problems():
            if vs and vg and name not in seen:
                seen.add(name)
                _CASES.append((name, scene, start, goal))
                if len(seen) >= limit:
                    break
    return _CASES


@pytest.mark.parametrize("batch", [1000, 100, 20, 5])
def test_fcit_restatement_equals_reference_fcit(batch):
    r = po.Ref("panda")
    deep = 0
    for name, scene, a, b in _cases():
        env = po.add_scene(po.RefEnv(), scene)
        want, want_cost, want_it = r.planner(2, env, a, b, batch_size=batch, max_samples=4000)
        got, got_cost, got_it = r.planner(3, env, a, b, batch_size=batch, max_samples=4000)
        assert got_it == want_it, (name, batch)
        assert got.shape == want.shape and np.array_equal(got, want), (name, batch)
        assert got_cost == want_cost, (name, batch)
        deep += want_it > 2
    assert batch == 1000 or deep > 0  # small batches really exercise the multi-iteration control flow


def test_reference_prm_roadmap_shape():
    # the compiled reference PRM over the exact NN stand-in (the source of tests/golden/planner.npz)
    r = po.Ref("panda")
    env = po.add_scene(po.RefEnv(), scenes.packed(scenes.sphere_cage()))
    V, E, it = r.planner(0, env, scenes.CAGE_A, scenes.CAGE_B, max_iterations=1500, max_samples=1500)
    assert it == 1501 and len(V) > 100 and np.allclose(V[0], scenes.CAGE_A, atol=1e-6) and np.allclose(V[1], scenes.CAGE_B, atol=1e-6)
    # adjacency is symmetric: every (i, j) entry has its (j, i) twin
    S = set(map(tuple, E.tolist()))
    assert all((j, i) in S for i, j in S)
