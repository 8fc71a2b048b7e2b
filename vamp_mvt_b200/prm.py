"""Batched PRM front-end: ``vamp.<robot>.prm`` / ``vamp.<robot>.roadmap`` of the reference
(bindings/robot_helper.hh:177-221, planning/prm.hh:43-300) over the C ABI's ``vmv_prm`` (csrc/vmv_planner.cu).

The reference draws one Halton sample at a time, checks it, asks its k-d tree for the PRM* neighbours and validates one
edge per neighbour.  Here the same roadmap comes out of three bulk steps on the GPU -- samples generated and validated
on the device, exact causal k-nearest neighbours by one kernel, all candidate edges as one indexed edge batch -- and a
host replay of union-find / A*.  Vertices, adjacency, iteration count and path are the reference's
(tests/test_planner.py compares them with the reference's own prm.hh compiled in place).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _lib
from .environment import Environment

# Robot::space_measure() of the reference (robots/panda.hh:111-114, ur5.hh:105-108, fetch.hh:117-120, baxter.hh:153-156):
# a float-returning function, so the literal is narrowed to float before PRMStarNeighborParams takes it as a double
SPACE_MEASURE = {"panda": 57376.4026747593, "ur5": 61528.90796697732, "fetch": 16384.87636281249, "baxter": 590532810.7756369}


@dataclass
class Roadmap:
    vertices: np.ndarray          # [n][dof]: start, goal, then the valid samples in stream order
    edges: np.ndarray             # [m][2] adjacency entries (from, to) in the reference's Roadmap::edges enumeration
    edge_cost: np.ndarray         # [m] the neighbour distance stored with each entry
    path: Optional[List[np.ndarray]] = None  # solve: start ... goal
    cost: float = float("inf")
    iterations: int = 0
    samples_drawn: int = 0
    edges_checked: int = 0
    stats: dict = field(default_factory=dict)

    def undirected_edges(self) -> np.ndarray:
        e = np.sort(self.edges.astype(np.int64), axis=1)
        return np.unique(e, axis=0)


def _run(robot, start, goal, environment: Optional[Environment], max_iterations: int, max_samples: int, solve: bool, fcit_batch: int = 0,
         optimize: bool = False) -> Roadmap:
    L = _lib.lib()
    env = environment if environment is not None else Environment()
    d = robot.dimension()
    s = _lib.f32(start).reshape(d)
    g = _lib.f32(goal).reshape(d)
    h = C.c_void_p()
    measure = float(np.float32(SPACE_MEASURE[robot.name]))
    if fcit_batch > 0:
        _lib.check(L.vmv_fcit(robot.id, env.handle, _lib.ptr(s), _lib.ptr(g), max_iterations, max_samples, fcit_batch, 1 if optimize else 0, C.byref(h)))
    else:
        _lib.check(L.vmv_prm(robot.id, env.handle, _lib.ptr(s), _lib.ptr(g), max_iterations, max_samples, measure, 1 if solve else 0, C.byref(h)))
    try:
        p = C.c_void_p()
        n = L.vmv_roadmap_vertices(h, C.byref(p))
        V = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n * d,)).reshape(n, d).copy() if n else np.zeros((0, d), np.float32)
        pe, pc = C.c_void_p(), C.c_void_p()
        m = L.vmv_roadmap_edges(h, C.byref(pe), C.byref(pc))
        E = np.ctypeslib.as_array(C.cast(pe, C.POINTER(C.c_uint32)), shape=(2 * m,)).reshape(m, 2).copy() if m else np.zeros((0, 2), np.uint32)
        W = np.ctypeslib.as_array(C.cast(pc, C.POINTER(C.c_float)), shape=(m,)).copy() if m else np.zeros(0, np.float32)
        pp, cost = C.c_void_p(), C.c_float(0)
        k = L.vmv_roadmap_path(h, C.byref(pp), C.byref(cost))
        path = None
        if k:
            P = np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_float)), shape=(k * d,)).reshape(k, d).copy()
            path = [P[i] for i in range(k)]
        ec = C.c_size_t(0)
        drawn = L.vmv_roadmap_work(h, C.byref(ec))
        return Roadmap(V, E, W, path, float(cost.value), int(L.vmv_roadmap_iterations(h)), int(drawn), int(ec.value))
    finally:
        L.vmv_roadmap_destroy(h)


def prm(robot, start, goal, environment: Optional[Environment] = None, max_iterations: int = 100000, max_samples: int = 100000) -> Roadmap:
    """vamp.<robot>.prm: PRM::solve with the Halton sampler and PRM* neighbour parameters."""
    return _run(robot, start, goal, environment, max_iterations, max_samples, True)


def fcit(robot, start, goal, environment: Optional[Environment] = None, max_iterations: int = 100000, max_samples: int = 100000,
         batch_size: int = 1000, optimize: bool = False) -> Roadmap:
    """vamp.<robot>.fcit: FCIT*::solve (planning/fcit.hh) -- host search, GPU sample batches and edge rows."""
    return _run(robot, start, goal, environment, max_iterations, max_samples, True, fcit_batch=batch_size, optimize=optimize)


def roadmap(robot, start, goal, environment: Optional[Environment] = None, max_iterations: int = 100000, max_samples: int = 100000) -> Roadmap:
    """vamp.<robot>.roadmap: PRM::build_roadmap."""
    return _run(robot, start, goal, environment, max_iterations, max_samples, False)
