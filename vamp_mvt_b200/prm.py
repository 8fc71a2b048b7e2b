"""Batched PRM front-end over the validation engine -- a first slice of SURVEY.md 8(f) rank 1, not a
port of the reference planner.

The reference's PRM (planning/prm.hh:43-196) is strictly serial: one sample, one ``fkcc``, its k nearest
roadmap vertices, one ``validate_motion`` per neighbour.  The work it generates is exactly what the
engine takes in bulk, so here a roadmap grows in ROUNDS: a batch of samples is validated in one call
(``vmv_validate_configs``), every valid sample proposes edges to its k nearest earlier vertices with the
reference's PRM* neighbour count k = ceil((e + e/d) ln n) (planning/roadmap.hh:49-56), all candidate edges
of the round are validated in one call as index pairs into the vertex table (``vmv_validate_edges_indexed``
-- 8 bytes per edge), and connectivity is tracked with union-find; start and goal are the first two
vertices and a shortest path is extracted when they meet.  Samples come from ``rng`` -- e.g. the
reference's Halton sequence, ``vamp.<robot>.halton()`` (vamp_mvt_b200/halton.py), drawn a batch at a time
-- or, without one, uniformly from a seeded generator; nearest neighbours come from scipy's k-d tree
(the reference's nigh tree is not restated) and vertices are connected in rounds -- so roadmaps are not
the reference's, but every edge in one is an edge the reference's ``validate_motion`` accepts.
"""
from __future__ import annotations

import heapq
import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .environment import Environment
from .halton import Halton


@dataclass
class Roadmap:
    vertices: np.ndarray
    edges: np.ndarray  # [m, 2] vertex indices of validated edges
    path: Optional[List[np.ndarray]] = None
    cost: float = float("inf")
    rounds: int = 0
    configs_checked: int = 0
    edges_checked: int = 0
    stats: dict = field(default_factory=dict)


def _find(parent, i):
    while parent[i] != i:
        parent[i] = parent[parent[i]]
        i = parent[i]
    return i


def prm(robot, start, goal, environment: Optional[Environment] = None, max_samples: int = 20000, batch: int = 4096,
        seed: int = 0, rng=None) -> Roadmap:
    from scipy.spatial import cKDTree

    d = robot.dimension()
    lo = np.asarray(robot.lower_bounds(), np.float32)
    hi = np.asarray(robot.upper_bounds(), np.float32)
    gen = np.random.default_rng(seed)
    start = np.asarray(start, np.float32).reshape(d)
    goal = np.asarray(goal, np.float32).reshape(d)
    rm = Roadmap(vertices=np.stack([start, goal]), edges=np.zeros((0, 2), np.int64))
    ends_ok = robot.validate_batch(rm.vertices, environment)
    rm.configs_checked += 2
    if not ends_ok.all():
        return rm
    # straight line first, as the reference does (prm.hh:57-70)
    rm.edges_checked += 1
    if robot.validate_motion(start, goal, environment):
        rm.edges = np.array([[0, 1]])
        rm.path, rm.cost = [start, goal], float(np.linalg.norm(goal - start))
        return rm
    parent = list(range(2))
    while len(rm.vertices) < max_samples:
        rm.rounds += 1
        if isinstance(rng, Halton) and rng.count + batch <= robot.halton_exact_limit():
            # samples generated and validated on the device; the valid ones are rebuilt here in closed form
            ok = robot.validate_halton(rng.count, batch, environment)
            new = rng.at(rng.count + np.nonzero(ok)[0])
            rng.advance(batch)
        else:
            if rng is not None:
                q = rng.take(batch)
            else:
                q = (lo + (hi - lo) * gen.random((batch, d), dtype=np.float32)).astype(np.float32)
            ok = robot.validate_batch(q, environment)
            new = q[ok]
        rm.configs_checked += batch
        if len(new) == 0:
            continue
        base = len(rm.vertices)
        V = np.vstack([rm.vertices, new]).astype(np.float32)
        n = len(V)
        k = int(math.ceil((math.e + math.e / d) * math.log(n)))
        tree = cKDTree(V)
        _, nbr = tree.query(new, k=min(k + 1, n))
        src = np.repeat(np.arange(base, n), nbr.shape[1])
        dst = nbr.reshape(-1)
        keep = dst < src  # an edge is proposed once, by its later vertex
        pairs = np.stack([src[keep], dst[keep]], axis=1).astype(np.uint32)
        valid = robot.validate_edges_indexed(V, pairs, environment)
        rm.edges_checked += len(pairs)
        good = pairs[valid].astype(np.int64)
        rm.vertices = V
        rm.edges = np.vstack([rm.edges, good])
        parent.extend(range(base, n))
        for a, b in good:
            ra, rb = _find(parent, int(a)), _find(parent, int(b))
            if ra != rb:
                parent[ra] = rb
        if _find(parent, 0) == _find(parent, 1):
            break
    if _find(parent, 0) != _find(parent, 1):
        return rm
    # Dijkstra on the validated edges
    adj = [[] for _ in range(len(rm.vertices))]
    w = np.linalg.norm(rm.vertices[rm.edges[:, 0]] - rm.vertices[rm.edges[:, 1]], axis=1)
    for (a, b), c in zip(rm.edges, w):
        adj[a].append((b, float(c)))
        adj[b].append((a, float(c)))
    dist = {0: 0.0}
    prev = {}
    heap = [(0.0, 0)]
    while heap:
        du, u = heapq.heappop(heap)
        if u == 1:
            break
        if du > dist.get(u, float("inf")):
            continue
        for v, c in adj[u]:
            if du + c < dist.get(v, float("inf")):
                dist[v] = du + c
                prev[v] = u
                heapq.heappush(heap, (du + c, v))
    node, idx = 1, [1]
    while node != 0:
        node = prev[node]
        idx.append(node)
    rm.path = [rm.vertices[i] for i in reversed(idx)]
    rm.cost = dist[1]
    return rm
