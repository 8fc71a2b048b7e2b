"""Generates tests/golden/planner.npz: roadmaps and solutions of the reference's OWN planners (planning/prm.hh,
fcit.hh compiled in place into oracle/_ref over the exact brute-force NN stand-in, with the reference's Halton
sampler) on the sphere cage of scripts/sphere_cage_example.py and on one MotionBenchMaker scene.  Run where
/root/reference is mounted; tests/test_planner.py compares the GPU front-end with the committed file."""
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from oracle import pyoracle as po  # noqa: E402
from tests import scenes  # noqa: E402
from tests.test_golden import mbm_problems  # noqa: E402


def main():
    po.build()
    r = po.Ref("panda")
    out = {}
    cases = {"cage": (scenes.packed(scenes.sphere_cage()), np.array(scenes.CAGE_A, np.float32), np.array(scenes.CAGE_B, np.float32))}
    for name, index, scene, start, goal, vs, vg, classic in mbm_problems():
        if name == "bookshelf_small" and vs and vg:
            cases["mbm_bookshelf_small"] = (scene, np.asarray(start, np.float32), np.asarray(goal, np.float32))
            break
    for key, (scene, a, b) in cases.items():
        env = po.add_scene(po.RefEnv(), scene)
        V, E, it = r.planner(0, env, a, b, max_iterations=4000, max_samples=4000)
        P, cost, it_s = r.planner(1, env, a, b)
        F, fcost, it_f = r.planner(2, env, a, b)
        for bs in (100, 20):
            Fb, cb, ib = r.planner(2, env, a, b, batch_size=bs, max_samples=4000)
            out.update({f"{key}_fcit{bs}_path": Fb, f"{key}_fcit{bs}_cost": cb, f"{key}_fcit{bs}_iterations": ib})
        out.update({f"{key}_start": a, f"{key}_goal": b, f"{key}_vertices": V, f"{key}_edges": E, f"{key}_iterations": it,
                    f"{key}_path": P, f"{key}_cost": cost, f"{key}_solve_iterations": it_s,
                    f"{key}_fcit_path": F, f"{key}_fcit_cost": fcost, f"{key}_fcit_iterations": it_f})
        for k in ("spheres", "cuboids", "capsules"):
            width = {"spheres": 4, "cuboids": 15, "capsules": 8}[k]
            out[f"{key}_{k}"] = np.asarray(scene.get(k, []), np.float32).reshape(len(scene.get(k, [])), width)
        kinds = {"spheres": 0, "cuboids": 1, "capsules": 2}
        order = scene.get("order") or ([("spheres", i) for i in range(len(scene.get("spheres", [])))] + [("cuboids", i) for i in range(len(scene.get("cuboids", [])))]
                                       + [("capsules", i) for i in range(len(scene.get("capsules", [])))])
        out[f"{key}_order"] = np.array([(kinds[k], i) for k, i in order], np.int32)
        print(key, "roadmap", V.shape, E.shape, it, "| solve", P.shape, cost, it_s, "| fcit", F.shape, fcost, it_f)
    out["cases"] = np.array(list(cases.keys()))
    path = REPO / "tests" / "golden" / "planner.npz"
    np.savez_compressed(path, **out)
    print("wrote", path, path.stat().st_size)


if __name__ == "__main__":
    main()
