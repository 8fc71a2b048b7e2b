"""``problem_dict_to_vamp`` (reference src/vamp/__init__.py:141-188): MotionBenchMaker problem
dictionaries -> Environment, with the reference's rules (cylinders become capsules, except in the
"box" problem where they become cuboids)."""
from __future__ import annotations

from typing import Any, Dict, List

from .environment import Environment
from .shapes import Cuboid, Cylinder, Sphere


def problem_dict_to_vamp(problem: Dict[str, Any], ignore_names: List[str] = []) -> Environment:
    env = Environment()
    for obj in problem.get("sphere", []):
        if obj["name"] not in ignore_names:
            env.add_sphere(Sphere(obj["position"], obj["radius"], name=obj["name"]))

    if problem.get("problem") == "box":
        for obj in problem.get("cylinder", []):
            if obj["name"] in ignore_names:
                continue
            env.add_cuboid(
                Cuboid(
                    obj["position"],
                    obj["orientation_euler_xyz"],
                    [obj["radius"], obj["radius"], obj["length"] / 2],
                    name=obj["name"],
                )
            )
    else:
        for obj in problem.get("cylinder", []):
            if obj["name"] in ignore_names:
                continue
            env.add_capsule(
                Cylinder(obj["position"], obj["orientation_euler_xyz"], obj["radius"], obj["length"], name=obj["name"])
            )

    for obj in problem.get("box", []):
        if obj["name"] not in ignore_names:
            env.add_cuboid(Cuboid(obj["position"], obj["orientation_euler_xyz"], obj["half_extents"], name=obj["name"]))

    return env
