#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the reference's own code (oracle/_ref/libvamp_ref.so, the
reference headers compiled in place by oracle/ref/Makefile).  Run in the container where
/root/reference is mounted; the fixtures are committed, the script documents how they were made.

Each fixture holds the inputs (configurations, edges, the scene as packed shape fields in insertion
order) and the reference's outputs (sphere_fk, validate verdicts, validate_motion verdicts)."""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import numpy as np

from oracle import pyoracle as po
from tests import scenes

KEEP_OUT = {"panda": 0.0, "ur5": 0.0, "fetch": 0.45, "baxter": 0.5}


def scene_arrays(sc):
    p = scenes.packed(sc)
    kinds = {"spheres": 0, "cuboids": 1, "capsules": 2}
    return dict(
        spheres=np.array(p["spheres"], np.float32).reshape(-1, 4),
        cuboids=np.array(p["cuboids"], np.float32).reshape(-1, 15),
        capsules=np.array(p["capsules"], np.float32).reshape(-1, 8),
        order=np.array([[kinds[k], i] for k, i in p["order"]], np.int32).reshape(-1, 2),
    )


def main():
    po.build()
    assert po.ref_available(), "oracle/_ref must be built (needs /root/reference)"
    out = REPO / "tests" / "golden"
    out.mkdir(parents=True, exist_ok=True)
    for robot in ["panda", "ur5", "fetch", "baxter"]:
        ref = po.Ref(robot)
        probe = scenes.random_configs(robot, 512, seed=1)

        def pick(seed):
            # first seeded random scene (from `seed` upwards) in which 10-90 % of random configurations are valid
            while True:
                sc = scenes.random_scene(seed, keep_out=KEEP_OUT[robot])
                frac = ref.validate_configs(po.add_scene(po.RefEnv(), scenes.packed(sc)), probe).mean()
                if 0.1 < frac < 0.9:
                    return sc
                seed += 1

        named = {"random0": pick(100), "random1": pick(200), "empty": {"order": []}}
        if robot == "panda":
            named.update(cage=scenes.sphere_cage(), table=scenes.table_shelf_scene(), box=scenes.box_scene())
        q = scenes.random_configs(robot, 256, seed=900)
        # a few configurations outside the joint limits as well
        q[:32] += np.random.default_rng(1).uniform(-0.7, 0.7, size=(32, q.shape[1])).astype(np.float32)
        a, b = scenes.random_edges(robot, 96, seed=901)
        data = dict(q=q, a=a, b=b, fk=ref.sphere_fk(q[:64]), scene_names=np.array(sorted(named)))
        for name in sorted(named):
            sc = named[name]
            env = po.add_scene(po.RefEnv(), scenes.packed(sc))
            for k, v in scene_arrays(sc).items():
                data[f"{name}_{k}"] = v
            data[f"{name}_valid"] = ref.validate_configs(env, q)
            data[f"{name}_edge_valid"] = ref.validate_edges(env, a, b)
        np.savez_compressed(out / f"{robot}.npz", **data)
        print(robot, {n: (float(data[f'{n}_valid'].mean()), float(data[f'{n}_edge_valid'].mean())) for n in sorted(named)})
    # the sphere cage known answers (reference scripts/sphere_cage_example.py:10-31,67)
    ref = po.Ref("panda")
    env = po.add_scene(po.RefEnv(), scenes.packed(scenes.sphere_cage()))
    qa = np.array([scenes.CAGE_A, scenes.CAGE_B], np.float32)
    np.savez_compressed(
        out / "sphere_cage.npz",
        a=qa[0], b=qa[1], validate=ref.validate_configs(env, qa), validate_motion=ref.validate_edges(env, qa[:1], qa[1:]),
        fk_a=ref.sphere_fk(qa[:1])[0],
    )


if __name__ == "__main__":
    main()
