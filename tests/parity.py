"""Parity accounting shared by the GPU tests, ``__graft_entry__.smoke()`` and ``bench.py``.

Bar (BASELINE.json north_star): verdicts bit-exact, except for units whose minimum signed clearance lies
within 1e-5 m of zero.  Every mismatching unit is therefore classified, never just counted:

* configuration: ``|min signed clearance| <= BAND`` (oracle ``or_clearance`` out[0]: f64 distances from the
  f32 FK centres over every fine / attachment sphere vs every obstacle and every allowed sphere pair);
* edge: some state of its rake schedule is inside the band (a mismatch needs one state on which the two
  sides disagree) -- ``or_edge_clearance``;
* environments with a pointcloud: the hierarchy is not conservative there (link bounding spheres wider than
  the tree's r_max are queried as they are), so a bounding-sphere query can decide a verdict
  by itself; such a unit is in-band when its nearest DECISION boundary (out[1], which adds bounding sphere
  vs cloud point pairs) is within the band.
"""
from __future__ import annotations

import numpy as np

BAND = 1e-5  # metres


def classify(clear: np.ndarray, has_cloud: bool) -> np.ndarray:
    """clear: [k][2] from Oracle.clearance / edge_clearance / point_clearance -> bool[k] in-band."""
    clear = np.asarray(clear, dtype=np.float64).reshape(-1, 2)
    ok = np.abs(clear[:, 0]) <= BAND
    if has_cloud:
        ok |= clear[:, 1] <= BAND
    return ok


def config_report(oracle, oenv, q, got, want, has_cloud: bool = False) -> dict:
    bad = np.nonzero(np.asarray(got, bool) != np.asarray(want, bool))[0]
    clear = oracle.clearance(oenv, q[bad]) if len(bad) else np.zeros((0, 2))
    ok = classify(clear, has_cloud)
    return {
        "units": int(len(q)),
        "mismatch_in_band": int(ok.sum()),
        "mismatch_outside": int((~ok).sum()),
        "outside": [(int(i), clear[k].tolist()) for k, i in enumerate(bad) if not ok[k]][:8],
    }


def edge_report(oracle, oenv, a, b, got, want, has_cloud: bool = False) -> dict:
    bad = np.nonzero(np.asarray(got, bool) != np.asarray(want, bool))[0]
    clear = oracle.edge_clearance(oenv, a[bad], b[bad]) if len(bad) else np.zeros((0, 2))
    ok = classify(clear, has_cloud)
    return {
        "units": int(len(a)),
        "mismatch_in_band": int(ok.sum()),
        "mismatch_outside": int((~ok).sum()),
        "outside": [(int(i), clear[k].tolist()) for k, i in enumerate(bad) if not ok[k]][:8],
    }


def assert_configs(oracle, oenv, q, got, want, what: str, has_cloud: bool = False) -> int:
    rep = config_report(oracle, oenv, q, got, want, has_cloud)
    assert rep["mismatch_outside"] == 0, f"{oracle.robot} {what}: config mismatches outside the {BAND} m band: {rep}"
    return rep["mismatch_in_band"]


def assert_edges(oracle, oenv, a, b, got, want, what: str, has_cloud: bool = False) -> int:
    rep = edge_report(oracle, oenv, a, b, got, want, has_cloud)
    assert rep["mismatch_outside"] == 0, f"{oracle.robot} {what}: edge mismatches outside the {BAND} m band: {rep}"
    return rep["mismatch_in_band"]
