"""Per-robot modules mirroring ``vamp.<robot>`` for the validation path (reference
bindings/robot_helper.hh:325-597): ``validate``, ``fk``, ``debug``, ``eefk`` for one configuration
plus the batched entry points the planners hand their work to.
"""
from __future__ import annotations

import ctypes as C
import json
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .environment import Environment
from .shapes import Sphere

_EMPTY_ENV: Optional[Environment] = None


def _empty_env() -> Environment:
    global _EMPTY_ENV
    if _EMPTY_ENV is None:
        _EMPTY_ENV = Environment()
    return _EMPTY_ENV


class Robot:
    def __init__(self, name: str):
        self.name = name
        self._L = _lib.lib()
        self.id = _lib.check(self._L.vmv_robot_id(name.encode()))
        self._dof = self._L.vmv_robot_dof(self.id)
        self._n_spheres = self._L.vmv_robot_n_spheres(self.id)
        self._resolution = self._L.vmv_robot_resolution(self.id)
        lo = np.zeros(self._dof, np.float32)
        rg = np.zeros(self._dof, np.float32)
        _lib.check(self._L.vmv_robot_bounds(self.id, _lib.ptr(lo), _lib.ptr(rg)))
        self._lower, self._range = lo, rg
        self._model = None

    # -- constants (robot_helper.hh:327-345) ------------------------------------------------
    def dimension(self) -> int:
        return self._dof

    def resolution(self) -> int:
        return self._resolution

    def n_spheres(self) -> int:
        return self._n_spheres

    def lower_bounds(self) -> List[float]:
        return self._lower.tolist()

    def upper_bounds(self) -> List[float]:
        return (self._lower + self._range).tolist()

    @property
    def model(self) -> dict:
        if self._model is None:
            path = Path(__file__).resolve().parent / "robots" / f"{self.name}.json"
            self._model = json.loads(path.read_text())
        return self._model

    def joint_names(self) -> List[str]:
        return list(self.model["joint_names"])

    def end_effector(self) -> str:
        return self.model["end_effector"]["link"]

    def min_max_radii(self):
        return (self.model["min_radius"], self.model["max_radius"])

    def _cfg(self, q) -> np.ndarray:
        a = _lib.f32(q).reshape(-1)
        if a.size != self._dof:
            raise ValueError(f"{self.name}: configuration must have {self._dof} values, got {a.size}")
        return a

    def in_bounds(self, q) -> bool:
        # Helper::validate's bound test (robot_helper.hh:258-262): descale, 0 <= v <= 1
        a = self._cfg(q)
        with np.errstate(divide="ignore", invalid="ignore"):
            d = (a - self._lower) * (np.float32(1) / self._range)
        return bool(np.all(d <= 1.0) and np.all(d >= 0.0))

    # -- single-unit API, names as in the reference -----------------------------------------
    def validate(self, configuration, environment: Optional[Environment] = None, check_bounds: bool = False) -> bool:
        """vamp.<robot>.validate (robot_helper.hh:255-267, 548-553)."""
        q = self._cfg(configuration)
        if check_bounds and not self.in_bounds(q):
            return False
        return bool(self.validate_batch(q[None, :], environment)[0])

    def validate_motion(self, a, b, environment: Optional[Environment] = None) -> bool:
        """validate_motion<Robot,8,resolution>(a,b,env) (planning/validate.hh:70-77)."""
        return bool(self.validate_motion_batch(self._cfg(a)[None, :], self._cfg(b)[None, :], environment)[0])

    def fk(self, configuration) -> List[Sphere]:
        """vamp.<robot>.fk (robot_helper.hh:234-247): the fine collision spheres."""
        out = self.fk_batch(self._cfg(configuration)[None, :])[0]
        return [Sphere(s[:3], s[3]) for s in out]

    def eefk(self, configuration) -> np.ndarray:
        """vamp.<robot>.eefk (robot_helper.hh:276-279): 4x4 end-effector frame.  Host-side chain
        product over the robot model with libm sin/cos, like the reference's scalar eefk."""
        q = self._cfg(configuration).astype(np.float64)
        m = self.model
        frames = []
        for b in m["bodies"]:
            if b["parent"] < 0:
                frames.append(np.eye(4))
                continue
            pre = np.vstack([np.array(b["T_pre"]), [0, 0, 0, 1]])
            ax = np.array(b["axis"], float)
            ax /= np.linalg.norm(ax)
            J = np.eye(4)
            v = q[b["dof"]]
            if b["jtype"] == "revolute":
                K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
                J[:3, :3] = np.eye(3) + np.sin(v) * K + (1 - np.cos(v)) * (K @ K)
            else:
                J[:3, 3] = ax * v
            frames.append(frames[b["parent"]] @ pre @ J)
        ee = m["end_effector"]
        T = frames[ee["body"]] @ np.vstack([np.array(ee["T"]), [0, 0, 0, 1]])
        return T.astype(np.float32)

    def debug(self, configuration, environment: Optional[Environment] = None):
        """vamp.<robot>.debug (robot_helper.hh:249-253): ([names hit per fine sphere], [(i, j) self pairs])."""
        env = environment if environment is not None else _empty_env()
        q = self._cfg(configuration)
        cap = 1 << 16
        eh = np.zeros((cap, 2), np.int32)
        sh = np.zeros((cap, 2), np.int32)
        ne, ns = C.c_size_t(0), C.c_size_t(0)
        _lib.check(
            self._L.vmv_debug(
                self.id, env.handle, _lib.ptr(q), _lib.ptr(eh), cap, C.byref(ne), _lib.ptr(sh), cap, C.byref(ns)
            )
        )
        per_sphere = [[] for _ in range(self._n_spheres)]
        for s, obj in eh[: ne.value]:
            name = env.names[obj] if obj < len(env.names) else ""
            per_sphere[s].append(name if name else str(obj))
        return per_sphere, [tuple(map(int, p)) for p in sh[: ns.value]]

    # -- batched API (host buffers; copies included) -----------------------------------------
    def validate_batch(self, configurations, environment: Optional[Environment] = None) -> np.ndarray:
        env = environment if environment is not None else _empty_env()
        q = _lib.f32(configurations).reshape(-1, self._dof)
        n = len(q)
        words = np.zeros((n + 31) // 32, np.uint32)
        _lib.check(self._L.vmv_validate_configs(self.id, env.handle, _lib.ptr(q), n, _lib.ptr(words)))
        return _lib.unpack_bits(words, n)

    def validate_motion_batch(self, a, b, environment: Optional[Environment] = None, resolution: int = 0) -> np.ndarray:
        env = environment if environment is not None else _empty_env()
        a = _lib.f32(a).reshape(-1, self._dof)
        b = _lib.f32(b).reshape(-1, self._dof)
        if a.shape != b.shape:
            raise ValueError("a and b must have the same shape")
        n = len(a)
        words = np.zeros((n + 31) // 32, np.uint32)
        _lib.check(
            self._L.vmv_validate_edges(self.id, env.handle, _lib.ptr(a), _lib.ptr(b), n, resolution, _lib.ptr(words))
        )
        return _lib.unpack_bits(words, n)

    def validate_edges_indexed(self, vertices, pairs, environment: Optional[Environment] = None, resolution: int = 0) -> np.ndarray:
        """Edge i joins vertices pairs[i, 0] and pairs[i, 1] of the vertex table (PRM-style edge sets,
        reference planning/prm.hh:136-146): 8 bytes per edge instead of two configurations."""
        env = environment if environment is not None else _empty_env()
        V = _lib.f32(vertices).reshape(-1, self._dof)
        P = np.ascontiguousarray(np.asarray(pairs, dtype=np.uint32).reshape(-1, 2))
        words = np.zeros((len(P) + 31) // 32, np.uint32)
        _lib.check(self._L.vmv_validate_edges_indexed(self.id, env.handle, _lib.ptr(V), len(V), _lib.ptr(P), len(P), resolution, _lib.ptr(words)))
        return _lib.unpack_bits(words, len(P))

    def filter_self_from_pointcloud(self, pointcloud, point_radius: float, configuration,
                                    environment: Optional[Environment] = None) -> np.ndarray:
        """Points of the cloud that neither overlap the robot at `configuration` nor collide with the
        environment, in their original order (reference bindings/robot_helper.hh:284-322, 555-561)."""
        env = environment if environment is not None else _empty_env()
        p = _lib.f32(pointcloud).reshape(-1, 3)
        q = self._cfg(configuration)
        words = np.zeros((len(p) + 31) // 32, np.uint32)
        _lib.check(self._L.vmv_filter_self_from_pointcloud(self.id, env.handle, _lib.ptr(q), _lib.ptr(p), len(p),
                                                           float(point_radius), _lib.ptr(words)))
        return p[_lib.unpack_bits(words, len(p))]

    def prm(self, start, goal, environment: Optional[Environment] = None, **kwargs):
        """``vamp.<robot>.prm`` (PRM::solve, reference planning/prm.hh:43-196) as bulk GPU steps: vamp_mvt_b200/prm.py."""
        from .prm import prm

        return prm(self, start, goal, environment, **kwargs)

    def fcit(self, start, goal, environment: Optional[Environment] = None, **kwargs):
        """``vamp.<robot>.fcit`` (FCIT*::solve, reference planning/fcit.hh:82-360): vamp_mvt_b200/prm.py."""
        from .prm import fcit

        return fcit(self, start, goal, environment, **kwargs)

    def roadmap(self, start, goal, environment: Optional[Environment] = None, **kwargs):
        """``vamp.<robot>.roadmap`` (PRM::build_roadmap, reference planning/prm.hh:198-300): vamp_mvt_b200/prm.py."""
        from .prm import roadmap

        return roadmap(self, start, goal, environment, **kwargs)

    def Path(self, waypoints=()):
        """``vamp.<robot>.Path`` (reference planning/plan.hh:10-169, bindings/robot_helper.hh:411-466)."""
        from .path import Path

        return Path(self, waypoints)

    def simplify(self, path, environment: Optional[Environment] = None, settings=None, rng=None):
        """``vamp.<robot>.simplify(path, environment, settings, rng)`` (reference bindings/robot_helper.hh:269-277,
        planning/simplify.hh:191-258) with the candidate edges of every routine validated as batches."""
        from .simplify import simplify

        return simplify(self, path, environment, settings, rng)

    def validate_halton(self, first: int, n: int, environment: Optional[Environment] = None, return_configs: bool = False):
        """Samples first .. first+n-1 of the reference's Halton sequence (random/halton.hh) generated and
        validated on the GPU (``vmv_validate_halton``): only one bit per sample crosses PCIe.  Returns the
        verdicts (and the configurations if asked)."""
        env = environment if environment is not None else _empty_env()
        words = np.zeros((n + 31) // 32, np.uint32)
        q = np.zeros((n, self._dof), np.float32) if return_configs else None
        _lib.check(self._L.vmv_validate_halton(self.id, env.handle, int(first), n, _lib.ptr(words),
                                               _lib.ptr(q) if return_configs else None))
        ok = _lib.unpack_bits(words, n)
        return (ok, q) if return_configs else ok

    def halton_exact_limit(self) -> int:
        return int(self._L.vmv_halton_exact_limit(self.id))

    def halton(self):
        """``vamp.<robot>.halton()`` (reference random/halton.hh): the deterministic configuration sampler."""
        from .halton import Halton

        return Halton(self)

    def fk_batch(self, configurations) -> np.ndarray:
        q = _lib.f32(configurations).reshape(-1, self._dof)
        out = np.zeros((len(q), self._n_spheres, 4), np.float32)
        _lib.check(self._L.vmv_sphere_fk(self.id, _lib.ptr(q), len(q), _lib.ptr(out)))
        return out
