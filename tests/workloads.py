"""The BASELINE.json workloads beyond C2/C3's plain batches, shared by the GPU tests and bench.py:

  C4  Fetch and UR5 FK+CC against a CAPT pointcloud of 100 k synthetic surface points plus a 256 x 256
      heightfield (r_point = vamp.constants.POINT_RADIUS, r_min / r_max = the robot's sphere radii)
  C5  PRM-style roadmap edges: (u32, u32) index pairs into a table of valid vertices, neighbours by construction
      (clusters of jittered configurations), generated per shard so that every rank can build its own slice.
"""
from __future__ import annotations

import numpy as np

from tests import scenes

C4_KEEP_OUT = {"fetch": 0.55, "ur5": 0.3, "panda": 0.3}


def synth_pointcloud(n: int, keep_out: float, seed: int = 0) -> np.ndarray:
    """Surface samples of a synthetic table / wall / object / floor scene, in the spirit of src/vamp/pointcloud.py."""
    rng = np.random.default_rng(seed)
    parts = [
        rng.uniform([0.35, -0.8, 0.38], [1.1, 0.8, 0.40], size=(n * 4 // 10, 3)),    # table top
        rng.uniform([-0.4, 0.75, 0.0], [0.6, 0.78, 1.4], size=(n * 3 // 10, 3)),      # wall / shelf back
        rng.normal([0.7, 0.2, 0.55], [0.06, 0.06, 0.1], size=(n * 2 // 10, 3)),      # object on the table
        rng.uniform([-1.2, -1.2, 0.0], [1.2, 1.2, 0.02], size=(n - n * 9 // 10, 3)),  # floor
    ]
    p = np.concatenate(parts).astype(np.float32)
    return p[np.hypot(p[:, 0], p[:, 1]) > keep_out]


def c4_heightfield(seed: int = 1):
    """256 x 256 cells of 2 cm, flat under the robot base (the robots stay well inside the field)."""
    import vamp_mvt_b200 as vmv

    rng = np.random.default_rng(seed)
    xd = yd = 256
    data = (0.15 * rng.random((yd, xd)) ** 4).astype(np.float32)
    yy, xx = np.mgrid[0:yd, 0:xd]
    data[np.hypot(xx - xd / 2, yy - yd / 2) < 40] = 0.0
    return vmv.make_heightfield([0, 0, -0.3], [0.02, 0.02, 1.0], [xd, yd], data), xd, yd, data


def c4_environment(robot: str, n_points: int = 100_000):
    """-> (product Environment, points, heightfield tuple, build nanoseconds)."""
    import vamp_mvt_b200 as vmv

    R = getattr(vmv, robot)
    rmin, rmax = R.min_max_radii()
    pts = synth_pointcloud(n_points, C4_KEEP_OUT[robot])
    env = vmv.Environment()
    build_ns = env.add_capt_pointcloud(pts, rmin, rmax, vmv.POINT_RADIUS)
    hf = c4_heightfield()
    env.add_heightfield(hf[0])
    return env, pts, hf, build_ns


def c4_checker_env(env_cls, robot: str, pts, hf, raw_only: bool = False):
    """The same environment for a CPU checker (po.RefEnv / po.OracleEnv); raw_only: the oracle's clearance
    accounting alone (raw cloud + heightfield, no CAPT build)."""
    import vamp_mvt_b200 as vmv

    R = getattr(vmv, robot)
    rmin, rmax = R.min_max_radii()
    e = env_cls()
    if raw_only:
        e.add_raw_cloud(pts, vmv.POINT_RADIUS)
    else:
        e.add_capt(pts, rmin, rmax, vmv.POINT_RADIUS)
    e.add_heightfield(hf[0].packed(), hf[1], hf[2], hf[3].reshape(-1))
    return e


# ---- C5 ---------------------------------------------------------------------------------------------
C5_CLUSTERS, C5_PER = 1 << 14, 64


def c5_vertices(seed: int = 5):
    """2^14 clusters x 64 jittered configurations (PRM neighbourhoods), clipped to the joint bounds: [2^20][7]."""
    import vamp_mvt_b200 as vmv

    rng = np.random.default_rng(seed)
    centres = scenes.random_configs("panda", C5_CLUSTERS, seed=seed)
    V = (centres[:, None, :] + rng.normal(0, 0.12, size=(C5_CLUSTERS, C5_PER, 7)).astype(np.float32)).reshape(-1, 7)
    lo = np.array(vmv.panda.lower_bounds(), np.float32)
    hi = np.array(vmv.panda.upper_bounds(), np.float32)
    return np.clip(V, lo, hi).astype(np.float32)


def c5_cluster_tables(valid: np.ndarray):
    """valid: bool[2^20] verdicts of the vertex table -> (clusters with >= 2 valid members, per cluster the
    count and the padded list of valid member slots)."""
    v = valid.reshape(C5_CLUSTERS, C5_PER)
    cnt = v.sum(axis=1).astype(np.int64)
    order = np.argsort(~v, axis=1, kind="stable").astype(np.int64)  # valid members first, in slot order
    good = np.nonzero(cnt >= 2)[0].astype(np.int64)
    return good, cnt, order


def c5_pairs_numpy(good, cnt, order, first: int, n: int, seed: int = 5) -> np.ndarray:
    """Edges [first, first + n) of the global list: both end points valid members of one cluster.  Generated in
    blocks of 2^16 edges seeded by the block index, so any shard can be produced on its own."""
    out = np.empty((n, 2), np.uint32)
    blk = 1 << 16
    b0, b1 = first // blk, (first + n + blk - 1) // blk
    at = 0
    for b in range(b0, b1):
        rng = np.random.default_rng([seed, b])
        c = good[rng.integers(0, len(good), size=blk)]
        k = cnt[c]
        i = (rng.random(blk) * k).astype(np.int64)
        j = (i + 1 + (rng.random(blk) * (k - 1)).astype(np.int64)) % k
        p = np.stack([c * C5_PER + order[c, i], c * C5_PER + order[c, j]], axis=1).astype(np.uint32)
        lo = max(first - b * blk, 0)
        hi = min(first + n - b * blk, blk)
        out[at : at + hi - lo] = p[lo:hi]
        at += hi - lo
    return out
