"""``vamp.<robot>.halton()``: the reference's deterministic configuration sampler
(random/halton.hh:8-109, binding bindings/robot_helper.hh), restated in closed form.

The reference advances a numerator / denominator pair per joint in f32 (halton.hh:76-107).  Up to its
``max_iterations`` = 10^6 both stay integers below 2^24, so the recurrence is exact and sample i is the
radical inverse of i in the joint's base (3, 5, 7, 11, ...; bases rotate left every 10^6 samples,
halton.hh:51-57, 79-85) -- computed here for a whole batch at once, then ``n / d`` in f32 and
``Robot::scale_configuration`` (q * range + lower as one FMA) exactly like ``next()`` does.  Pinned
against the compiled reference in tests/test_simplify.py.
"""
from __future__ import annotations

import numpy as np

PRIMES = (3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59)  # halton.hh:17-33
MAX_ITERATIONS = 1000000  # halton.hh:12


class Halton:
    """RNG<Robot> with the reference's ``next()`` / ``reset()`` plus batched ``take(n)``.  ``dist`` is the
    reference's integer/real distribution (random/distribution.hh)."""

    def __init__(self, robot):
        from .simplify import Distribution

        self.dim = robot.dimension()
        if self.dim > len(PRIMES):
            raise ValueError("Halton: at most 16 joints")
        self._lower = np.asarray(robot.lower_bounds(), np.float32).astype(np.float64)
        self._range = np.asarray(robot._range, np.float32).astype(np.float64)
        self.dist = Distribution()
        self._replay_cache = {}
        self.reset()

    def reset(self) -> None:
        self.count = 0  # samples drawn so far
        self.dist.reset()

    def _unit(self, index: np.ndarray) -> np.ndarray:
        """index: 0-based sample numbers -> [k][dim] radical inverses (f32 n / f32 d)."""
        # the first epoch lasts max_iterations samples, every later one max_iterations + 1: the call that
        # trips the limit resets the counter to 0 and already returns the new epoch's first sample
        # (halton.hh:78-85)
        later = np.maximum(index - MAX_ITERATIONS, 0)
        first = index < MAX_ITERATIONS
        epoch = np.where(first, 0, 1 + later // (MAX_ITERATIONS + 1))
        i = np.where(first, index + 1, later % (MAX_ITERATIONS + 1) + 1).astype(np.int64)
        out = np.zeros((len(index), self.dim), np.float32)
        for j in range(self.dim):
            # after e rotations joint j uses the base that started at position (j + e) mod dim
            base = np.asarray(PRIMES, np.int64)[(j + epoch) % self.dim]
            num = np.zeros(len(index), np.int64)
            den = np.ones(len(index), np.int64)
            rest = i.copy()
            while (rest > 0).any():
                live = rest > 0
                num = np.where(live, num * base + rest % base, num)
                den = np.where(live, den * base, den)
                rest = np.where(live, rest // base, 0)
            out[:, j] = num.astype(np.float32) / den.astype(np.float32)
            # bases 29 and 31 (joints 9, 10: Baxter only) pass 2^24 at i = base^4 < 10^6: from there on the
            # reference's f32 recurrence is no longer the radical inverse ("numerical precision degrades",
            # halton.hh:11) -- replay it step by step for those samples
            inexact = (den >= (1 << 24))
            if inexact.any():
                out[inexact, j] = self._replay(i[inexact], base[inexact])
        return out

    def _replay(self, i: np.ndarray, base: np.ndarray) -> np.ndarray:
        """The f32 recurrence of halton.hh:87-103 itself, from the last exact state (i = base^4 - 1)."""
        res = np.zeros(len(i), np.float32)
        f = np.float32
        for b in np.unique(base):
            sel = np.nonzero(base == b)[0]
            top = int(i[sel].max())
            start = int(b) ** 4 - 1  # all digits b-1: n = d - 1, d = b^4
            key = int(b)
            cache = self._replay_cache.setdefault(key, {"at": start, "n": f(start), "d": f(int(b) ** 4), "vals": {}})
            bf = f(b)
            n, d, at, vals = cache["n"], cache["d"], cache["at"], cache["vals"]
            while at < top:
                x = f(d - n)
                if x == f(1):
                    d = f(np.floor(f(d * bf)))
                    n = f(1)
                else:
                    y = f(np.floor(f(d / bf)))
                    while x <= y:
                        y = f(np.floor(f(y / bf)))
                    n = f(f(np.floor(f(f(bf + f(1)) * y))) - x)
                at += 1
                vals[at] = f(n / d)
            cache.update(n=n, d=d, at=at)
            res[sel] = [vals[int(k)] for k in i[sel]]
        return res

    def take(self, n: int) -> np.ndarray:
        """The next n samples, scaled to the joint ranges: what n calls of ``next()`` return."""
        idx = self.count + np.arange(n, dtype=np.int64)
        self.count += n
        u = self._unit(idx).astype(np.float64)
        return (u * self._range + self._lower).astype(np.float32)

    def at(self, index) -> np.ndarray:
        """Samples with the given 0-based numbers (what the device generated for them)."""
        u = self._unit(np.asarray(index, np.int64).reshape(-1)).astype(np.float64)
        return (u * self._range + self._lower).astype(np.float32)

    # StreamRNG protocol of vamp_mvt_b200.simplify (peek / advance / next)
    def peek(self, k: int) -> np.ndarray:
        u = self._unit(self.count + np.arange(k, dtype=np.int64)).astype(np.float64)
        return (u * self._range + self._lower).astype(np.float32)

    def advance(self, k: int) -> None:
        self.count += k

    def next(self) -> np.ndarray:
        return self.take(1)[0]

    @property
    def samples(self):  # perturb_path only asks whether a stream exists
        return self
