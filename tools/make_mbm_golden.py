#!/usr/bin/env python3
"""Generates tests/golden/mbm_panda.npz: the MotionBenchMaker problems the reference ships
(resources/panda/problems.tar.bz2) with the REFERENCE's own verdicts for every start and goal.

Runs in the container where /root/reference is mounted.  The MoveIt YAML is converted with the
reference's own rules: resources/problem_tar_to_pkl_json.py:30-84 (scene / request parsing, using the
reference's vendored src/vamp/transformations.py, imported here) and src/vamp/__init__.py:141-188
(problem_dict_to_vamp, restated in vamp_mvt_b200/problems.py).  Verdicts come from the reference's
headers compiled in place (oracle/_ref/libvamp_ref.so).  resources/README.md:146 publishes
"Solved / Valid / Total: 699 / 699 / 700" for the seven classic problem sets; the script asserts it."""
import importlib.util
import re
import sys
import tarfile
from collections import defaultdict
from pathlib import Path

import numpy as np
import yaml

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from oracle import pyoracle as po
from vamp_mvt_b200.shapes import Cuboid, Cylinder, Sphere

REF = Path("/root/reference")
CLASSIC = ["bookshelf_small", "bookshelf_tall", "bookshelf_thin", "box", "cage", "table_pick", "table_under_pick"]


def load_transformations():
    spec = importlib.util.spec_from_file_location("ref_transformations", REF / "src/vamp/transformations.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    tfm = load_transformations()

    def transform_from_yaml(tf):
        return np.dot(tfm.translation_matrix(tf["position"]), tfm.quaternion_matrix(tf["orientation"]))

    robot = "panda"
    joints = [f"panda_joint{i}" for i in range(1, 8)]
    scenes, requests = defaultdict(dict), defaultdict(dict)
    tar = tarfile.open(REF / "resources" / robot / "problems.tar.bz2", "r:bz2")
    for member in tar.getmembers():
        if not member.isfile():
            continue
        _, problem, filename = member.name.split("/")
        problem = problem.replace(f"_{robot}", "")
        data = yaml.load(tar.extractfile(member).read(), Loader=getattr(yaml, "CLoader", yaml.Loader))
        index = int(re.findall(r"\d+", filename)[0])
        if "scene" in filename:
            objs = []  # (kind, shape) in problem_dict_to_vamp order: spheres, cylinders, boxes
            by_kind = {"sphere": [], "cylinder": [], "box": []}
            for co in data["world"]["collision_objects"]:
                base = tfm.identity_matrix() if "pose" not in co else transform_from_yaml(co["pose"])
                prim = co["primitives"][0]
                pose = np.dot(base, transform_from_yaml(co["primitive_poses"][0]))
                pos = tfm.translation_from_matrix(pose).tolist()
                euler = list(tfm.euler_from_matrix(pose))
                by_kind[prim["type"]].append((pos, euler, prim["dimensions"], co["id"]))
            for pos, euler, dim, name in by_kind["sphere"]:
                objs.append(("sphere", Sphere(pos, dim[0], name=name)))
            for pos, euler, dim, name in by_kind["cylinder"]:
                length, radius = dim[0], dim[1]
                if problem == "box":  # src/vamp/__init__.py:153-171
                    objs.append(("cuboid", Cuboid(pos, euler, [radius, radius, length / 2], name=name)))
                else:
                    objs.append(("capsule", Cylinder(pos, euler, radius, length, name=name)))
            for pos, euler, dim, name in by_kind["box"]:
                objs.append(("cuboid", Cuboid(pos, euler, [d / 2 for d in dim], name=name)))
            scenes[problem][index] = objs
        else:
            js = data["start_state"]["joint_state"]
            start = [js["position"][js["name"].index(j)] for j in joints]
            jc = data["goal_constraints"][0]["joint_constraints"]
            names, pos = [e["joint_name"] for e in jc], [e["position"] for e in jc]
            goal = [pos[names.index(j)] for j in joints]
            requests[problem][index] = (start, goal)

    ref = po.Ref(robot)
    sets = sorted(scenes)
    kinds = {"sphere": 0, "cuboid": 1, "capsule": 2}
    spheres, cuboids, capsules, order, offsets = [], [], [], [], [0]
    starts, goals, v_start, v_goal, set_id, index = [], [], [], [], [], []
    for si, name in enumerate(sets):
        for idx in sorted(scenes[name]):
            if idx not in requests[name]:
                continue
            env = po.RefEnv()
            for kind, shape in scenes[name][idx]:
                f = np.asarray(shape.packed(), np.float32)
                if kind == "sphere":
                    env.add_sphere(f), order.append((0, len(spheres))), spheres.append(f)
                elif kind == "cuboid":
                    env.add_cuboid(f), order.append((1, len(cuboids))), cuboids.append(f)
                else:
                    env.add_capsule(f), order.append((2, len(capsules))), capsules.append(f)
            offsets.append(len(order))
            s, g = requests[name][idx]
            q = np.array([s, g], np.float32)
            v = ref.validate_configs(env, q)
            starts.append(q[0]), goals.append(q[1]), v_start.append(v[0]), v_goal.append(v[1])
            set_id.append(si), index.append(idx)
    v_start, v_goal, set_id = np.array(v_start), np.array(v_goal), np.array(set_id)
    valid = v_start & v_goal
    classic = np.isin(set_id, [sets.index(n) for n in CLASSIC])
    print("problems", len(valid), "valid", int(valid.sum()), "| classic sets:", int(classic.sum()), "valid", int(valid[classic].sum()))
    for si, name in enumerate(sets):
        print(f"  {name:28s} {int(valid[set_id == si].sum())}/{int((set_id == si).sum())}")
    assert int(classic.sum()) == 700 and int(valid[classic].sum()) == 699, "resources/README.md:146 publishes 699 / 700"
    out = REPO / "tests" / "golden" / "mbm_panda.npz"
    np.savez_compressed(
        out, sets=np.array(sets), set_id=set_id.astype(np.int16), index=np.array(index, np.int16),
        spheres=np.array(spheres, np.float32).reshape(-1, 4), cuboids=np.array(cuboids, np.float32).reshape(-1, 15),
        capsules=np.array(capsules, np.float32).reshape(-1, 8), order=np.array(order, np.int32).reshape(-1, 2),
        offsets=np.array(offsets, np.int32), start=np.array(starts, np.float32), goal=np.array(goals, np.float32),
        valid_start=v_start, valid_goal=v_goal, classic=classic)
    print("wrote", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
