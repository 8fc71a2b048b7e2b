"""``vamp.<robot>.simplify`` with its edge checks batched (reference planning/simplify.hh:14-258,
settings planning/simplify_settings.hh:7-56, binding bindings/robot_helper.hh:269-277).

The reference calls ``validate_motion`` once per candidate edge.  Here every routine collects the
candidates that do not depend on each other and hands them to the GPU as one edge batch
(``vmv_validate_edges`` / ``vmv_validate_edges_indexed``), then replays the reference's sequential
decisions on the verdict bits -- the resulting path is the reference's, waypoint for waypoint:

* ``shortcut_path`` (simplify.hh:116-141): the verdict of (i, j) only depends on the two original
  waypoints, so all n(n-1)/2 - (n-1) non-adjacent pairs go out as ONE indexed edge batch and the greedy
  scan runs over the bit matrix;
* ``smooth_bspline`` (simplify.hh:14-52): after ``subdivide`` the even waypoints move, the odd ones do
  not, so the candidates of one step are independent: two edges per candidate, one batch per step;
* ``perturb_path`` (simplify.hh:143-189): the ``perturbation_attempts`` candidates of one step are
  validated together (two edges each); samples are consumed exactly as the reference consumes them
  (up to and including the first accepted attempt);
* ``reduce_path_vertices`` (simplify.hh:54-114): each step depends on the previous one; one edge per call.

Random draws: ``Distribution`` restates the reference's ``vamp::rng::Distribution``
(random/distribution.hh:9-47: ``std::mt19937`` seeded 0 + libstdc++'s ``uniform_real_distribution<float>``);
the configuration sampler behind ``RNG::next()`` (Halton / xorshift in the reference, out of scope here)
is any callable returning unit-cube samples -- ``StreamRNG`` replays an array.

Arithmetic follows the reference's f32 vector code: ``interpolate`` = ``a + (b - a) * alpha``
(vector/interface.hh:422-425; GCC contracts it to one FMA under the reference's ``-ffp-contract=fast``,
restated as a f64 product-sum rounded once), ``distance`` = sqrt of the AVX ``hsum`` tree
(vector/interface.hh:397-420, vector/avx.hh:441-452).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

from .environment import Environment
from .path import Path

BSPLINE, REDUCE, SHORTCUT, PERTURB = 0, 1, 2, 3  # SimplifyRoutine (simplify_settings.hh:7-13)
_ROUTINE_IDS = {"BSPLINE": BSPLINE, "REDUCE": REDUCE, "SHORTCUT": SHORTCUT, "PERTURB": PERTURB}


@dataclass
class BSplineSettings:
    max_steps: int = 1
    min_change: float = 0.1
    midpoint_interpolation: float = 0.5


@dataclass
class ReduceSettings:
    max_steps: int = 10
    max_empty_steps: int = 5
    range_ratio: float = 0.5


@dataclass
class ShortcutSettings:
    pass


@dataclass
class PerturbSettings:
    max_steps: int = 10
    max_empty_steps: int = 5
    perturbation_attempts: int = 5
    range: float = 0.1


@dataclass
class SimplifySettings:
    max_iterations: int = 5
    interpolate: int = 0
    operations: List[int] = field(default_factory=lambda: [SHORTCUT, BSPLINE])
    reduce: ReduceSettings = field(default_factory=ReduceSettings)
    shortcut: ShortcutSettings = field(default_factory=ShortcutSettings)
    bspline: BSplineSettings = field(default_factory=BSplineSettings)
    perturb: PerturbSettings = field(default_factory=PerturbSettings)


@dataclass
class PlanningResult:
    """planning/plan.hh:171-179."""

    path: Path
    cost: float = 0.0
    nanoseconds: int = 0
    iterations: int = 0
    size: List[int] = field(default_factory=list)


# ---- reference arithmetic (host side, f32) ------------------------------------------------------
def interpolate(a: np.ndarray, b: np.ndarray, alpha: float) -> np.ndarray:
    """``a.interpolate(b, alpha)`` = a + (b - a) * alpha, the product-sum rounded once (FMA)."""
    a = np.asarray(a, np.float32)
    d = (np.asarray(b, np.float32) - a).astype(np.float64)
    return (d * np.float64(np.float32(alpha)) + a.astype(np.float64)).astype(np.float32)


def distance(a: np.ndarray, b: np.ndarray) -> np.float32:
    """``a.distance(b)``: rows of 8 summed elementwise, then the AVX hsum tree, then sqrt."""
    d = np.asarray(b, np.float32) - np.asarray(a, np.float32)
    r0, r1 = np.zeros(8, np.float32), np.zeros(8, np.float32)
    r0[: min(len(d), 8)] = d[:8]
    r1[: max(len(d) - 8, 0)] = d[8:16]
    # a second row (Baxter) enters as GCC contracts row0*row0 + row1*row1 in the reference build: the first
    # product fused onto the rounded second one
    lane = (r0.astype(np.float64) ** 2 + (r1 * r1).astype(np.float32).astype(np.float64)).astype(np.float32)
    s = (lane[4:] + lane[:4]).astype(np.float32)
    return np.sqrt(np.float32(np.float32(s[0] + s[2]) + np.float32(s[1] + s[3])))


class Distribution:
    """vamp::rng::Distribution (random/distribution.hh:9-47): mt19937(0) + uniform_real_distribution<float>."""

    def __init__(self):
        self.reset()

    def reset(self) -> None:
        # RandomState's legacy integer seeding is init_genrand(seed), which is std::mt19937::seed(seed)
        self._rs = np.random.RandomState(0)

    def _raw32(self) -> int:
        return int(self._rs.randint(0, 1 << 32, dtype=np.uint64))

    def uniform_01(self) -> np.float32:
        # libstdc++ generate_canonical<float, 24>: one 32-bit draw / 2^32 in float; 1.0 is pulled below 1
        u = np.float32(self._raw32()) / np.float32(4294967296.0)
        return np.nextafter(np.float32(1), np.float32(0)) if u >= np.float32(1) else u

    def uniform_real(self, low: float, high: float) -> np.float32:
        low, high = np.float32(low), np.float32(high)
        span = np.float64(np.float32(high - low))
        return np.float32(span * np.float64(self.uniform_01()) + np.float64(low))

    def uniform_integer(self, low: int, high: int) -> int:
        r = int(np.floor(self.uniform_real(np.float32(low), np.float32(np.float64(high) + 1.0))))
        return high if r > high else r


class StreamRNG:
    """RNG<Robot> (random/rng.hh:9-17) whose ``next()`` replays an array of unit-cube samples."""

    def __init__(self, samples=None, dim: int = 0):
        self.samples = None if samples is None else np.asarray(samples, np.float32).reshape(-1, dim or np.shape(samples)[-1])
        self.at = 0
        self.dist = Distribution()

    def reset(self) -> None:
        self.at = 0
        self.dist.reset()

    def peek(self, k: int) -> np.ndarray:
        n = len(self.samples)
        return self.samples[(self.at + np.arange(k)) % n]

    def advance(self, k: int) -> None:
        self.at += k

    def next(self) -> np.ndarray:
        s = self.peek(1)[0]
        self.advance(1)
        return s


EdgeValidator = Callable[[np.ndarray, np.ndarray], np.ndarray]


def _validator(robot, environment: Optional[Environment], validate_edges: Optional[EdgeValidator]) -> EdgeValidator:
    if validate_edges is not None:
        return validate_edges
    return lambda a, b: robot.validate_motion_batch(a, b, environment)


def _as_list(path) -> List[np.ndarray]:
    return [np.asarray(w, np.float32) for w in path]


# ---- routines -----------------------------------------------------------------------------------
def shortcut_path(robot, path: Path, environment: Optional[Environment] = None, settings=None,
                  validate_edges: Optional[EdgeValidator] = None,
                  validate_indexed: Optional[Callable[[np.ndarray, np.ndarray], np.ndarray]] = None) -> bool:
    """simplify.hh:116-141.  All non-adjacent waypoint pairs in one indexed edge batch; then the greedy
    scan (first i, farthest reachable j) on the bits."""
    n = len(path)
    if n < 3:
        return False
    P = np.stack(_as_list(path))
    iu, ju = np.triu_indices(n, k=2)
    pairs = np.stack([iu, ju], axis=1).astype(np.uint32)
    if validate_indexed is not None:
        ok = validate_indexed(P, pairs)
    elif validate_edges is not None:
        ok = validate_edges(P[iu], P[ju])
    else:
        ok = robot.validate_edges_indexed(P, pairs, environment)
    free = np.zeros((n, n), bool)
    free[iu, ju] = np.asarray(ok, bool)

    keep = list(range(n))  # indices of the original waypoints still in the path
    result = False
    i = 0
    while i + 2 < len(keep):
        for j in range(len(keep) - 1, i + 1, -1):
            if free[keep[i], keep[j]]:
                del keep[i + 1 : j]
                result = True
                break
        i += 1
    if result:
        path[:] = [P[k] for k in keep]
    return result


def smooth_bspline(robot, path: Path, environment: Optional[Environment] = None,
                   settings: Optional[BSplineSettings] = None, validate_edges: Optional[EdgeValidator] = None) -> bool:
    """simplify.hh:14-52.  One edge batch per step: (prev, midpoint) and (midpoint, next) for every even
    waypoint whose midpoint moved more than ``min_change``."""
    st = settings or BSplineSettings()
    check = _validator(robot, environment, validate_edges)
    if len(path) < 3:
        return False
    changed = False
    for _ in range(st.max_steps):
        path.subdivide()
        P = _as_list(path)
        cand, mids = [], []
        for index in range(2, len(P) - 1, 2):
            t1 = interpolate(P[index], P[index - 1], st.midpoint_interpolation)
            t2 = interpolate(P[index], P[index + 1], st.midpoint_interpolation)
            mid = interpolate(t1, t2, 0.5)
            if distance(P[index], mid) > np.float32(st.min_change):
                cand.append(index)
                mids.append(mid)
        updated = False
        if cand:
            M = np.stack(mids)
            A = np.concatenate([np.stack([P[i - 1] for i in cand]), M])
            B = np.concatenate([M, np.stack([P[i + 1] for i in cand])])
            ok = np.asarray(check(A, B), bool)
            k = len(cand)
            for c, index in enumerate(cand):
                if ok[c] and ok[k + c]:
                    path[index] = mids[c]
                    updated = True
        changed |= updated
        if not updated:
            break
    return changed


def reduce_path_vertices(robot, path: Path, environment: Optional[Environment] = None,
                         settings: Optional[ReduceSettings] = None, rng: Optional[StreamRNG] = None,
                         validate_edges: Optional[EdgeValidator] = None) -> bool:
    """simplify.hh:54-114 (sequential: every step sees the previous step's erasure)."""
    st = settings or ReduceSettings()
    check = _validator(robot, environment, validate_edges)
    rng = rng or StreamRNG()
    if len(path) < 3:
        return False
    max_steps = st.max_steps or len(path)
    max_empty = st.max_empty_steps or len(path)
    result = False

    def step() -> bool:
        initial = len(path)
        max_n = initial - 1
        rng_range = 1 + int(np.floor(np.float32(0.5) + np.float32(initial) * np.float32(st.range_ratio)))
        p0 = rng.dist.uniform_integer(0, max_n)
        p1 = rng.dist.uniform_integer(max(p0 - rng_range, 0), min(max_n, p0 + rng_range))
        if abs(p0 - p1) < 2:
            if p0 < max_n - 1:
                p1 = p0 + 2
            elif p0 > 1:
                p1 = p0 - 2
            else:
                return False
        if p0 > p1:
            p0, p1 = p1, p0
        if bool(np.asarray(check(path[p0][None, :], path[p1][None, :]))[0]):
            del path[p0 + 1 : p1]
            return True
        return False

    i = no_change = 0
    while i < max_steps or no_change < max_empty:
        if step():
            no_change = 0
            result = True
        i, no_change = i + 1, no_change + 1
    return result


def perturb_path(robot, path: Path, environment: Optional[Environment] = None,
                 settings: Optional[PerturbSettings] = None, rng: Optional[StreamRNG] = None,
                 validate_edges: Optional[EdgeValidator] = None) -> bool:
    """simplify.hh:143-189.  The attempts of one step are validated as one batch; the sample stream
    advances as in the reference (through the first accepted attempt)."""
    st = settings or PerturbSettings()
    check = _validator(robot, environment, validate_edges)
    if rng is None or rng.samples is None:
        raise ValueError("perturb_path needs an RNG with a sample stream")
    if len(path) < 3:
        return False
    lo = np.asarray(robot.lower_bounds(), np.float32).astype(np.float64)
    span = (np.asarray(robot.upper_bounds(), np.float32) - np.asarray(robot.lower_bounds(), np.float32)).astype(np.float64)
    max_steps = st.max_steps or len(path)
    max_empty = st.max_empty_steps or len(path)
    changed = False
    step = no_change = 0
    while step < max_steps and no_change < max_empty:
        idx = rng.dist.uniform_integer(1, len(path) - 2)
        state, before, after = path[idx], path[idx - 1], path[idx + 1]
        old_cost = np.float32(distance(before, state) + distance(after, state))
        k = st.perturbation_attempts
        # Robot::scale_configuration: q * s_m + s_a (robots/<robot>.hh), one FMA
        samples = (rng.peek(k).astype(np.float64) * span + lo).astype(np.float32)
        cands = [interpolate(state, s, st.range) for s in samples]
        better = [np.float32(distance(before, c) + distance(after, c)) < old_cost for c in cands]
        sel = [c for c in range(k) if better[c]]
        taken = k
        if sel:
            C_ = np.stack([cands[c] for c in sel])
            A = np.concatenate([np.repeat(before[None, :], len(sel), 0), np.repeat(after[None, :], len(sel), 0)])
            ok = np.asarray(check(A, np.concatenate([C_, C_])), bool)
            for s_i, c in enumerate(sel):
                if ok[s_i] and ok[len(sel) + s_i]:
                    no_change = 0
                    changed = True
                    path[idx] = cands[c]
                    taken = c + 1
                    break
        rng.advance(taken)
        step, no_change = step + 1, no_change + 1
    return changed


def interpolate_to_n_states(path: Path, n: int) -> None:
    """Path::interpolate_to_n_states (planning/plan.hh:51-110)."""
    n_p = len(path)
    if n_p < 2 or n < n_p:
        return
    P = _as_list(path)
    seg = [distance(P[i], P[i + 1]) for i in range(n_p - 1)]
    remaining = np.float32(0)
    for s in seg:
        remaining = np.float32(remaining + s)
    if remaining < np.finfo(np.float32).eps:
        return
    out = []
    n1 = n_p - 1
    for i in range(n1):
        a, b = P[i], P[i + 1]
        out.append(a)
        max_n_states = n + i - n_p
        if max_n_states > 0:
            if i + 1 == n1:
                ns = max_n_states + 2
            else:
                # 0.5 + n * segment / remaining: size_t * float -> float, + double
                ns = int(np.floor(0.5 + np.float64(np.float32(np.float32(n) * seg[i]) / remaining))) + 1
            ns = min(ns - 2, max_n_states) if ns > 2 else 0
            v = (b - a).astype(np.float32)
            for k in range(1, ns + 1):
                f = np.float32(np.float32(k) / np.float32(ns))
                out.append((np.float64(f) * v.astype(np.float64) + a.astype(np.float64)).astype(np.float32))
            n -= ns + 1
            remaining = np.float32(remaining - seg[i])
        else:
            n -= 1
    out.append(P[-1])
    path[:] = out


def simplify(robot, path: Sequence, environment: Optional[Environment] = None,
             settings: Optional[SimplifySettings] = None, rng: Optional[StreamRNG] = None,
             validate_edges: Optional[EdgeValidator] = None) -> PlanningResult:
    """simplify.hh:191-258."""
    t0 = time.perf_counter_ns()
    st = settings or SimplifySettings()
    check = _validator(robot, environment, validate_edges)
    src = _as_list(path)
    result = PlanningResult(path=Path(robot))

    def done():
        result.nanoseconds = time.perf_counter_ns() - t0
        result.cost = result.path.cost()
        return result

    if len(src) == 2 or (len(src) > 2 and bool(np.asarray(check(src[0][None, :], src[-1][None, :]))[0])):
        result.path[:] = [src[0], src[-1]]
        return done()

    result.path[:] = src
    if st.interpolate:
        interpolate_to_n_states(result.path, st.interpolate)

    ops = {
        BSPLINE: lambda: smooth_bspline(robot, result.path, environment, st.bspline, validate_edges),
        REDUCE: lambda: reduce_path_vertices(robot, result.path, environment, st.reduce, rng, validate_edges),
        SHORTCUT: lambda: shortcut_path(robot, result.path, environment, st.shortcut, validate_edges),
        PERTURB: lambda: perturb_path(robot, result.path, environment, st.perturb, rng, validate_edges),
    }
    if len(src) > 2:
        for _ in range(st.max_iterations):
            result.iterations += 1
            any_change = False
            for op in st.operations:
                any_change |= ops[_ROUTINE_IDS.get(op, op) if isinstance(op, str) else op]()
            if not any_change:
                break
    return done()
