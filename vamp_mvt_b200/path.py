"""``vamp.<robot>.Path`` for the validation path (reference planning/plan.hh:10-169): a list of
configurations with ``cost``, ``subdivide``, ``interpolate_to_resolution`` and ``validate``.  ``validate``
is the only member on the hot path: the reference loops ``validate_motion`` over consecutive waypoints
(plan.hh:155-168); here the segments go to the GPU as one edge batch."""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _lib
from .environment import Environment


class Path(list):
    def __init__(self, robot, waypoints=()):
        super().__init__(np.asarray(w, np.float32).reshape(robot.dimension()) for w in waypoints)
        self.robot = robot

    def numpy(self) -> np.ndarray:
        return np.stack(self).astype(np.float32) if len(self) else np.zeros((0, self.robot.dimension()), np.float32)

    def append(self, configuration) -> None:
        super().append(np.asarray(configuration, np.float32).reshape(self.robot.dimension()))

    def insert(self, i: int, configuration) -> None:
        super().insert(i, np.asarray(configuration, np.float32).reshape(self.robot.dimension()))

    def cost(self) -> float:
        """plan.hh:12-31: l2 segment lengths (the reference's f32 summation order) added up front to back;
        infinity for fewer than two waypoints."""
        from .simplify import distance

        if len(self) < 2:
            return float("inf")
        total = np.float32(0)
        for a, b in zip(self[:-1], self[1:]):
            total = np.float32(total + distance(a, b))
        return float(total)

    def subdivide(self) -> None:
        """plan.hh:33-49: insert the midpoint of every segment."""
        if len(self) < 2:
            return
        p = self.numpy()
        out = []
        for a, b in zip(p[:-1], p[1:]):
            out += [a, a + np.float32(0.5) * (b - a)]
        out.append(p[-1])
        self[:] = out

    def interpolate_to_resolution(self, resolution: int) -> None:
        """plan.hh:112-153."""
        if len(self) < 2:
            return
        from .simplify import distance, interpolate

        p = self.numpy()
        out = []
        for a, b in zip(p[:-1], p[1:]):
            seg = distance(a, b)
            n = int(np.float32(seg * np.float32(resolution)))
            out.append(a)
            if seg < np.float32(1.0) / np.float32(resolution):
                continue
            for i in range(1, n):
                out.append(interpolate(a, b, np.float32(i) / np.float32(n)))
        out.append(p[-1])
        self[:] = out

    def interpolate_to_n_states(self, n: int) -> None:
        """plan.hh:51-110."""
        from .simplify import interpolate_to_n_states

        interpolate_to_n_states(self, int(n))

    def validate(self, environment: Optional[Environment] = None) -> bool:
        """plan.hh:155-168: every consecutive pair passes validate_motion at the robot's resolution."""
        if len(self) < 2:
            return True
        p = self.numpy()
        return bool(self.robot.validate_motion_batch(p[:-1], p[1:], environment).all())
