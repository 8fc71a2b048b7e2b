"""vamp_mvt_b200 -- B200-native batched configuration / motion validation engine.

Drop-in for the validation path of the reference ``vamp`` package (src/vamp/__init__.py):
``Environment``, ``Sphere``, ``Cuboid``, ``Cylinder``, ``Attachment``, ``make_heightfield`` and the
per-robot modules ``panda``, ``ur5``, ``fetch``, ``baxter`` with ``validate``, ``fk``, ``debug``,
``eefk`` -- plus batched entry points (``validate_batch``, ``validate_motion_batch``, ``fk_batch``).

Everything computes on the GPU through the C ABI in include/vamp_b200.h (libvamp_b200.so, hand-written
sm_100a CUDA).  There is no CPU fallback: importing works anywhere, calling needs a B200.
"""
from .shapes import Attachment, Cuboid, Cylinder, HeightField, Sphere, make_heightfield  # noqa: F401
from .environment import Environment  # noqa: F401
from .robot import Robot  # noqa: F401
from .problems import problem_dict_to_vamp  # noqa: F401
from .pointcloud import filter_pointcloud, filter_pointcloud_centervox  # noqa: F401
from .simplify import (  # noqa: F401  (vamp.SimplifySettings / vamp.SimplifyRoutine, bindings/settings.cc)
    BSplineSettings, PerturbSettings, ReduceSettings, ShortcutSettings, SimplifySettings, StreamRNG,
)

ROBOT_NAMES = ("panda", "ur5", "fetch", "baxter")
robots = list(ROBOT_NAMES)

# vamp.constants.POINT_RADIUS (src/vamp/constants.py:25)
POINT_RADIUS = 0.0025

_instances = {}


def __getattr__(name):
    # lazily build vamp.<robot>-style modules so that importing the package never touches CUDA
    if name in ROBOT_NAMES:
        if name not in _instances:
            _instances[name] = Robot(name)
        return _instances[name]
    raise AttributeError(name)
