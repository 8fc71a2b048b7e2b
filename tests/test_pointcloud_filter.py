"""CenterVox pointcloud filter (reference collision/filter_centervox.hh): the CPU restatement
(oracle/vamp_oracle.c) against the reference's own code compiled in place, and the CUDA path against both."""
import ctypes as C

import numpy as np
import pytest

import vamp_mvt_b200 as vmv
from oracle import pyoracle as po

HAVE_REF = po.ref_available() and hasattr(po.ref_lib(), "ref_filter_centervox")


def cloud(kind: str, n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.uniform(-1.6, 1.6, size=(n, 3)).astype(np.float32)
    if kind == "surfaces":  # table + wall + blob, as the C4 cloud
        a = rng.uniform([0.3, -0.8, 0.38], [1.1, 0.8, 0.40], size=(n // 2, 3))
        b = rng.uniform([-0.4, 0.75, 0.0], [0.6, 0.78, 1.4], size=(n // 4, 3))
        c = rng.normal([0.7, 0.2, 0.55], [0.06, 0.06, 0.1], size=(n - n // 2 - n // 4, 3))
        p = np.concatenate([a, b, c])
        return p[rng.permutation(len(p))].astype(np.float32)
    if kind == "lattice":
        # coordinates on a coarse lattice, a few ulps of jitter on some: exact ties and near-ties of the
        # squared distances, duplicates -- the cases where rounding and insertion order decide
        # (a two-layer slab: the reference's pool only holds 5 % of the grid, filter_centervox.hh:112-118)
        p = (rng.integers(-40, 41, size=(n, 3)) / 32.0).astype(np.float32)
        p[:, 2] = rng.integers(0, 2, size=n) / np.float32(32.0)
        j = rng.integers(-2, 3, size=(n, 3))
        p = np.where(rng.random((n, 3)) < 0.3, np.nextafter(p, p + np.sign(j).astype(np.float32)), p).astype(np.float32)
        return p
    raise ValueError(kind)


CASES = [
    # kind, n, voxel, range, origin, ws_min, ws_max
    ("uniform", 2000, 0.08, 1.4, [0, 0, 0.3], [-1.4, -1.4, -1.1], [1.4, 1.4, 1.7]),
    ("surfaces", 60000, 0.02, 1.19, [0, 0, 0.333], [-1.19, -1.19, -0.857], [1.19, 1.19, 1.523]),
    ("surfaces", 20000, 0.005, 1.19, [0, 0, 0.333], [-1.19, -1.19, -0.857], [1.19, 1.19, 1.523]),  # grid capped at 255
    ("lattice", 50000, 0.1, 1.3, [0.01, 0, 0], [-1.25, -1.25, -1.25], [1.25, 1.25, 1.25]),
    ("lattice", 50000, 0.0625, 2.5, [0, 0, 0], [-1.25, -1.25, -1.25], [1.25, 1.25, 1.25]),
    ("lattice", 30000, 0.07, 0.9, [0.2, -0.1, 0.05], [-1.0, -0.5, -1.25], [1.25, 1.0, 0.75]),  # anisotropic box
    ("uniform", 300, 0.1, 5.0, [0, 0, 0], [-1.6, -1.6, -1.6], [1.6, 1.6, 1.6]),
    ("uniform", 40000, 0.08, 1.4, [0, 0, 0.3], [-1.4, -1.4, -1.1], [1.4, 1.4, 1.7]),  # pool exhausted: the reference throws
]


def with_specials(p: np.ndarray, seed: int) -> np.ndarray:
    p = p.copy()
    rng = np.random.default_rng(seed)
    k = rng.choice(len(p), 12, replace=False)
    p[k[0]] = [np.nan, 0.1, 0.2]
    p[k[1]] = [0.1, np.nan, np.nan]
    p[k[2]] = [np.inf, 0, 0]
    p[k[3]] = [-np.inf, 0, 0]
    p[k[4:8]] = p[k[8:12]]  # exact duplicates
    return p


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (or built without the filter)")
@pytest.mark.parametrize("case", range(len(CASES)))
def test_oracle_filter_matches_compiled_reference(case):
    kind, n, vox, rng_, origin, lo, hi = CASES[case]
    for seed in range(3):
        p = with_specials(cloud(kind, n, seed), seed) if seed else cloud(kind, n, seed)
        want = po.ref_filter_centervox(p, vox, rng_, origin, lo, hi)
        got = po.filter_centervox(p, vox, rng_, origin, lo, hi)
        if want is None:
            assert got is None
            continue
        assert got is not None and len(got) == len(want) and len(want) > 0
        assert np.array_equal(p[got], want, equal_nan=True)  # same points, same order


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
def test_oracle_filter_edge_cases():
    lo, hi = [-1, -1, -1], [1, 1, 1]
    assert len(po.filter_centervox(np.zeros((0, 3), np.float32), 0.1, 1.0, [0, 0, 0], lo, hi)) == 0
    # everything culled by range / by the box
    p = cloud("uniform", 2000, 0)
    for args in ((0.1, 1e-3, [5, 5, 5], lo, hi), (0.1, 10.0, [0, 0, 0], [3, 3, 3], [4, 4, 4])):
        assert len(po.ref_filter_centervox(p, *args)) == 0 and len(po.filter_centervox(p, *args)) == 0
    # pool exhausted: (width / voxel)^3 * 0.05 voxels only (filter_centervox.hh:112-118) -> the reference throws
    dense = cloud("uniform", 20000, 1) * np.float32(0.6)
    assert po.ref_filter_centervox(dense, 0.2, 5.0, [0, 0, 0], lo, hi) is None
    assert po.filter_centervox(dense, 0.2, 5.0, [0, 0, 0], lo, hi) is None


def test_abi_argument_errors_without_gpu():
    L = vmv._lib.lib()
    k = C.c_size_t(7)
    idx = np.zeros(8, np.uint32)
    z = np.zeros(3, np.float32)
    o = np.ones(3, np.float32)
    p = np.zeros((4, 3), np.float32)
    ptr = vmv._lib.ptr
    assert L.vmv_filter_pointcloud_centervox(ptr(p), 0, 0.1, 1.0, ptr(z), ptr(z), ptr(o), ptr(idx), 8, C.byref(k)) == 0 and k.value == 0
    assert L.vmv_filter_pointcloud_centervox(ptr(p), 4, 0.0, 1.0, ptr(z), ptr(z), ptr(o), ptr(idx), 8, C.byref(k)) != 0
    assert L.vmv_filter_pointcloud_centervox(ptr(p), 4, 0.1, 1.0, ptr(z), ptr(o), ptr(z), ptr(idx), 8, C.byref(k)) != 0
    assert L.vmv_filter_pointcloud_centervox(None, 4, 0.1, 1.0, ptr(z), ptr(z), ptr(o), ptr(idx), 8, C.byref(k)) != 0
    with pytest.raises(NotImplementedError):
        vmv.filter_pointcloud(p, 0.01, 1.0, 0.1, z, z, o, True, "scdf")


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(CASES)))
def test_gpu_filter_matches_reference_and_oracle(case):
    kind, n, vox, rng_, origin, lo, hi = CASES[case]
    for seed in range(3):
        p = with_specials(cloud(kind, n, seed), seed) if seed else cloud(kind, n, seed)
        want_idx = po.filter_centervox(p, vox, rng_, origin, lo, hi)
        if want_idx is None:
            with pytest.raises(Exception):
                vmv.filter_pointcloud_centervox(p, vox, rng_, origin, lo, hi)
            continue
        got_idx = vmv.filter_pointcloud_centervox(p, vox, rng_, origin, lo, hi, return_indices=True)
        assert np.array_equal(got_idx, want_idx.astype(np.int64))
        if HAVE_REF:
            got, ns = vmv.filter_pointcloud(p, 0.0, rng_, vox, origin, lo, hi, True, "centervox")
            assert np.array_equal(got, po.ref_filter_centervox(p, vox, rng_, origin, lo, hi), equal_nan=True) and ns > 0


@pytest.mark.gpu
def test_gpu_filter_large_cloud_properties():
    # 2e6 points: beyond what the CPU checkers are asked to do in the suite; size-independent properties
    rng = np.random.default_rng(5)
    p = cloud("surfaces", 2_000_000, 9)
    vox, origin, lo, hi = 0.03, [0, 0, 0.333], [-1.19, -1.19, -0.857], [1.19, 1.19, 1.523]
    idx = vmv.filter_pointcloud_centervox(p, vox, 1.19, origin, lo, hi, return_indices=True)
    assert 1000 < len(idx) <= 32768 and len(np.unique(idx)) == len(idx)
    kept = p[idx]
    # one point per voxel
    inv = np.float32(min(255, int(np.ceil(np.float32(2.38) / np.float32(vox)))) / np.float32(2.38))
    v = np.clip(((kept - np.asarray(lo, np.float32)) * inv).astype(np.int32), 0, 254)
    assert len(np.unique(v, axis=0)) == len(kept)
    # idempotent on its own output (each kept point is alone in its voxel), and invariant to a permutation
    # of the input up to order
    again = vmv.filter_pointcloud_centervox(kept, vox, 1.19, origin, lo, hi)
    assert np.array_equal(again, kept)
    perm = rng.permutation(len(p))
    idx2 = vmv.filter_pointcloud_centervox(p[perm], vox, 1.19, origin, lo, hi, return_indices=True)
    a = {tuple(r) for r in np.round(kept, 6)}
    b = {tuple(r) for r in np.round(p[perm][idx2], 6)}
    assert len(a ^ b) <= 4  # only exact distance ties may resolve differently under another insertion order
    # agrees with the CPU restatement on the full cloud
    assert np.array_equal(idx, po.filter_centervox(p, vox, 1.19, origin, lo, hi).astype(np.int64))
