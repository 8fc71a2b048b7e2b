#!/usr/bin/env python3
"""Secondary workloads of BASELINE.json, measured on one GPU (device-resident inputs, CUDA events via
the library's own stream sync + host timer around >= 5 back-to-back launches):

  C4  Fetch and UR5 FK+CC vs a CAPT pointcloud of 100k synthetic points plus a 256x256 heightfield
  C5  PRM-style roadmap edge validation, 10^8 Panda edges as (u32,u32) index pairs into a vertex table

Usage: bench_extra.py [c4] [c5] [--edges N]
       python -m torch.distributed.run --nproc-per-node N ... bench_extra.py c5 --edges N
           C5 sharded over N GPUs (contiguous word-aligned shards, one NCCL all-gather of the
           verdict bitmask inside the timed region, device timing, max over ranks)
Prints one JSON line per workload."""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

import vamp_mvt_b200 as vmv
from tests import scenes
from vamp_mvt_b200 import _lib

L = _lib.lib()


def synth_pointcloud(n, keep_out, seed=0):
    """Surface samples of a synthetic table/shelf/wall scene, in the spirit of src/vamp/pointcloud.py."""
    rng = np.random.default_rng(seed)
    parts = [
        rng.uniform([0.35, -0.8, 0.38], [1.1, 0.8, 0.40], size=(n * 4 // 10, 3)),   # table top
        rng.uniform([-0.4, 0.75, 0.0], [0.6, 0.78, 1.4], size=(n * 3 // 10, 3)),     # wall / shelf back
        rng.normal([0.7, 0.2, 0.55], [0.06, 0.06, 0.1], size=(n * 2 // 10, 3)),     # object on the table
        rng.uniform([-1.2, -1.2, 0.0], [1.2, 1.2, 0.02], size=(n - n * 9 // 10, 3)),  # floor
    ]
    p = np.concatenate(parts).astype(np.float32)
    return p[np.hypot(p[:, 0], p[:, 1]) > keep_out]


def time_launches(fn, reps=5):
    fn()
    _lib.check(L.vmv_stream_sync(None))
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    _lib.check(L.vmv_stream_sync(None))
    return (time.perf_counter() - t0) / reps


def c4():
    for robot, keep in (("fetch", 0.55), ("ur5", 0.3)):
        R = getattr(vmv, robot)
        rmin, rmax = R.min_max_radii()
        pts = synth_pointcloud(100_000, keep)
        env = vmv.Environment()
        build_ns = env.add_capt_pointcloud(pts, rmin, rmax, vmv.POINT_RADIUS)
        rng = np.random.default_rng(1)
        xd = yd = 256
        data = (0.15 * rng.random((yd, xd)) ** 4).astype(np.float32)
        yy, xx = np.mgrid[0:yd, 0:xd]
        data[np.hypot(xx - xd / 2, yy - yd / 2) < 40] = 0.0
        hf = vmv.make_heightfield([0, 0, -0.3], [0.02, 0.02, 1.0], [xd, yd], data)
        env.add_heightfield(hf)
        n = 1 << 18
        q = scenes.random_configs(robot, n, seed=0)
        dq, db = L.vmv_dev_alloc(q.nbytes), L.vmv_dev_alloc((n + 31) // 32 * 4)
        _lib.check(L.vmv_memcpy_h2d(dq, _lib.ptr(q), q.nbytes, None))
        h = env.handle
        t = time_launches(lambda: _lib.check(L.vmv_validate_configs_dev(R.id, h, dq, n, db, None)))
        words = np.zeros((n + 31) // 32, np.uint32)
        _lib.check(L.vmv_memcpy_d2h(_lib.ptr(words), db, words.nbytes, None))
        _lib.check(L.vmv_stream_sync(None))
        out = {"workload": f"C4 {robot}: 2^18 configs vs CAPT({len(pts)} pts) + 256x256 heightfield",
               "value": n / t, "unit": "configs/s", "ms": t * 1e3, "capt_build_ms": build_ns / 1e6,
               "valid_fraction": float(_lib.unpack_bits(words, n).mean())}
        if "--cpu" in sys.argv:
            # the reference's own CAPT + validate on the host cores, on a bounded sample of the same batch
            from oracle import pyoracle as po
            if po.ref_available():
                renv = po.RefEnv()
                t0 = time.perf_counter()
                renv.add_capt(pts, rmin, rmax, vmv.POINT_RADIUS)
                out["reference_capt_build_ms"] = (time.perf_counter() - t0) * 1e3
                renv.add_heightfield(hf.packed(), xd, yd, data.reshape(-1))
                ns = 1 << 15
                thr = po.host_threads()
                sec = po.Ref(robot).time_configs(renv, q[:ns], thr, reps=2)
                want = po.Ref(robot).validate_configs(renv, q[:ns], threads=thr)
                out["cpu_reference"] = {"value": ns / sec, "unit": "configs/s", "cores": thr, "sample": f"first 2^15 configs of the batch"}
                out["mismatches_vs_reference_on_sample"] = int((want != _lib.unpack_bits(words, n)[:ns]).sum())
        print(json.dumps(out), flush=True)
        L.vmv_dev_free(dq), L.vmv_dev_free(db)


def filter_bench(n=4_000_000):
    """CenterVox filter (reference collision/filter_centervox.hh): n surface points, 2 cm voxels, the
    reference's MBM-style arguments (src/vamp/pointcloud.py:150-174).  Points per second with the cloud
    resident in HBM and through the host-pointer call, beside the reference's own filter on one host core
    (it is a sequential walk)."""
    import ctypes as C
    rng = np.random.default_rng(0)
    a = rng.uniform([0.3, -0.8, 0.38], [1.1, 0.8, 0.40], size=(n // 2, 3))
    b = rng.uniform([-0.4, 0.75, 0.0], [0.6, 0.78, 1.4], size=(n // 4, 3))
    c = rng.normal([0.7, 0.2, 0.55], [0.06, 0.06, 0.1], size=(n - n // 2 - n // 4, 3))
    p = np.concatenate([a, b, c])
    p = np.ascontiguousarray(p[rng.permutation(len(p))].astype(np.float32))
    vox, rad = 0.02, 1.19
    origin = np.array([0, 0, 0.333], np.float32)
    lo, hi = origin - np.float32(rad), origin + np.float32(rad)
    idx = np.zeros(32768, np.uint32)
    k = C.c_size_t(0)
    dp = L.vmv_dev_alloc(p.nbytes)
    _lib.check(L.vmv_memcpy_h2d(dp, _lib.ptr(p), p.nbytes, None))
    _lib.check(L.vmv_stream_sync(None))
    args = (len(p), vox, rad, _lib.ptr(origin), _lib.ptr(lo), _lib.ptr(hi), _lib.ptr(idx), len(idx), C.byref(k))

    def wall(fn, reps=10):
        for _ in range(2):
            fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps

    t_dev = wall(lambda: _lib.check(L.vmv_filter_pointcloud_centervox_dev(dp, *args)))
    kept_dev = idx[: k.value].copy()
    t_host = wall(lambda: _lib.check(L.vmv_filter_pointcloud_centervox(_lib.ptr(p), *args)))
    assert np.array_equal(kept_dev, idx[: k.value])
    out = {"workload": f"CenterVox filter: {len(p)} surface points, 2 cm voxels, {k.value} kept",
           "resident": {"points_per_s": len(p) / t_dev, "ms": t_dev * 1e3, "GBps_algorithmic": len(p) * 12 / t_dev / 1e9,
                        "note": "call with device-resident points: table reset + insert + compact + D2H of the kept records + host ordering"},
           "e2e": {"points_per_s": len(p) / t_host, "ms": t_host * 1e3, "h2d_bytes": int(p.nbytes)}}
    from oracle import pyoracle as po
    if po.ref_available() and hasattr(po.ref_lib(), "ref_filter_centervox"):
        ns = min(len(p), 1_000_000)
        t0 = time.perf_counter()
        want = po.ref_filter_centervox(p[:ns], vox, rad, origin, lo, hi)
        sec = time.perf_counter() - t0
        got = vmv.filter_pointcloud_centervox(p[:ns], vox, rad, origin, lo, hi)
        out["cpu_reference"] = {"points_per_s": ns / sec, "cores": 1, "sample": f"first {ns} points (incl. the harness's copy into std::vector<Point>)"}
        out["identical_to_reference_on_sample"] = bool(np.array_equal(got, want))
    print(json.dumps(out), flush=True)
    L.vmv_dev_free(dp)


def c5_inputs(n_edges):
    R = vmv.panda
    env = scenes.build_product_env(scenes.table_shelf_scene())
    rng = np.random.default_rng(5)
    # vertex table: clusters of nearby valid configurations (PRM neighbourhoods)
    n_clusters, per = 1 << 14, 64
    centres = scenes.random_configs("panda", n_clusters, seed=5)
    V = (centres[:, None, :] + rng.normal(0, 0.12, size=(n_clusters, per, 7)).astype(np.float32)).reshape(-1, 7)
    lo = np.array(R.lower_bounds(), np.float32)
    hi = np.array(R.upper_bounds(), np.float32)
    V = np.clip(V, lo, hi).astype(np.float32)
    valid = R.validate_batch(V, env)
    # edges: pairs inside a cluster whose two vertices are valid
    c = rng.integers(0, n_clusters, size=n_edges, dtype=np.int64)
    i = rng.integers(0, per, size=n_edges, dtype=np.int64)
    j = (i + rng.integers(1, per, size=n_edges, dtype=np.int64)) % per
    pairs = np.stack([c * per + i, c * per + j], axis=1).astype(np.uint32)
    keep = valid[pairs[:, 0]] & valid[pairs[:, 1]]
    return env, V, np.ascontiguousarray(pairs[keep])


def c5(n_edges):
    import os

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        return c5_sharded(n_edges)
    R = vmv.panda
    env, V, pairs = c5_inputs(n_edges)
    n = len(pairs)
    dV, dP, dB = L.vmv_dev_alloc(V.nbytes), L.vmv_dev_alloc(pairs.nbytes), L.vmv_dev_alloc((n + 31) // 32 * 4)
    _lib.check(L.vmv_memcpy_h2d(dV, _lib.ptr(V), V.nbytes, None))
    _lib.check(L.vmv_memcpy_h2d(dP, _lib.ptr(pairs), pairs.nbytes, None))
    h = env.handle
    t = time_launches(lambda: _lib.check(L.vmv_validate_edges_indexed_dev(R.id, h, dV, len(V), dP, n, 0, dB, None)), reps=3)
    words = np.zeros((n + 31) // 32, np.uint32)
    _lib.check(L.vmv_memcpy_d2h(_lib.ptr(words), dB, words.nbytes, None))
    _lib.check(L.vmv_stream_sync(None))
    d = np.linalg.norm(V[pairs[:100000, 0]] - V[pairs[:100000, 1]], axis=1)
    print(json.dumps({"workload": f"C5: {n} PRM-style Panda edges (index pairs into {len(V)} vertices), table/shelf scene",
                      "value": n / t, "unit": "edges/s", "ms": t * 1e3, "n_gpus": 1, "mean_edge_length_rad": float(d.mean()),
                      "valid_fraction": float(_lib.unpack_bits(words, n).mean())}), flush=True)


def c5_sharded(n_edges):
    """Every rank builds the same roadmap (seeded), validates its contiguous word-aligned shard of
    the edge list and takes part in ONE all-gather of verdict words (NCCL); strong scaling."""
    import os

    import torch
    import torch.distributed as dist

    from vamp_mvt_b200 import sharding

    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.check(L.vmv_set_device(local))
    R = vmv.panda
    env, V, pairs = c5_inputs(n_edges)
    n = len(pairs)
    lo, hi = sharding.shard_bounds(n, rank, world)
    per = sharding.words_per_rank(n, world)
    dV = torch.from_numpy(V).cuda()
    dP = torch.from_numpy(pairs[lo:hi].view(np.int32)).cuda()
    local_words = torch.zeros(per, dtype=torch.int32, device="cuda")
    gathered = torch.zeros(per * world, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    h = env.handle

    def step():
        _lib.check(L.vmv_validate_edges_indexed_dev(R.id, h, dV.data_ptr(), len(V), dP.data_ptr(), hi - lo, 0, local_words.data_ptr(), stream))
        dist.all_gather_into_tensor(gathered, local_words)

    for _ in range(2):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        words = gathered.cpu().numpy().view(np.uint32)[: (n + 31) // 32]
        print(json.dumps({"workload": f"C5: {n} PRM-style Panda edges (index pairs into {len(V)} vertices), table/shelf scene",
                          "value": n / (ms * 1e-3), "unit": "edges/s", "ms": ms, "n_gpus": world, "scaling": "strong",
                          "collective": f"one NCCL all_gather of {per * 4} bytes per rank per step",
                          "valid_fraction": float(_lib.unpack_bits(words, n).mean())}), flush=True)
    dist.destroy_process_group()


def robots():
    """All four robots on one primitives scene (15 objects): device-resident configs and edges."""
    for robot, keep in (("panda", 0.0), ("ur5", 0.0), ("fetch", 0.45), ("baxter", 0.5)):
        R = getattr(vmv, robot)
        env = scenes.build_product_env(scenes.random_scene(2, keep_out=keep))
        h = env.handle
        n, ne = 1 << 20, 1 << 17
        q = scenes.random_configs(robot, n, seed=0)
        a, b = scenes.random_edges(robot, ne, seed=0)
        dq, db = L.vmv_dev_alloc(q.nbytes), L.vmv_dev_alloc((n + 31) // 32 * 4)
        da, dbb = L.vmv_dev_alloc(a.nbytes), L.vmv_dev_alloc(b.nbytes)
        for d, hst in ((dq, q), (da, a), (dbb, b)):
            _lib.check(L.vmv_memcpy_h2d(d, _lib.ptr(hst), hst.nbytes, None))
        tc = time_launches(lambda: _lib.check(L.vmv_validate_configs_dev(R.id, h, dq, n, db, None)), reps=10)
        words = np.zeros((n + 31) // 32, np.uint32)
        _lib.check(L.vmv_memcpy_d2h(_lib.ptr(words), db, words.nbytes, None))
        _lib.check(L.vmv_stream_sync(None))
        vc = float(_lib.unpack_bits(words, n).mean())
        te = time_launches(lambda: _lib.check(L.vmv_validate_edges_dev(R.id, h, da, dbb, ne, 0, db, None)), reps=5)
        print(json.dumps({"workload": f"{robot}: 2^20 configs / 2^17 edges vs random scene (15 primitives)",
                          "configs_per_s": n / tc, "edges_per_s": ne / te, "valid_fraction": vc}), flush=True)
        for d in (dq, db, da, dbb):
            L.vmv_dev_free(d)


if __name__ == "__main__":
    what = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c4", "c5"]
    if "robots" in what:
        robots()
    n_edges = int(sys.argv[sys.argv.index("--edges") + 1]) if "--edges" in sys.argv else 100_000_000
    if "filter" in what:
        filter_bench()
    if "c4" in what:
        c4()
    if "c5" in what:
        c5(n_edges)
