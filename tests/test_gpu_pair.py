"""GPU parity tests of the two-view initialisation: argus_pair_triangulate through the C ABI and the
multiTriangulator built on it, against golden output of the reference's own triangulate.py, against OpenCV called live
(fundamental matrix), against the CPU restatement on seeded scenes, and through geometric properties at a large size."""
import os
import numpy as np
import pytest

from argus_gui_b200 import scenes, triangulate
from oracle import tri_port

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dense(s, m, n, seed, drop=0.1):
    dense = np.full((n, 2 * m), np.nan); uv = dense.reshape(n, m, 2); uv[s["vmask"] != 0] = s["x"]
    rng = np.random.default_rng(seed); d = rng.random((n, m)) < drop; d[:, -1] = False; uv[d] = np.nan
    return dense


def test_multitriangulator_golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "tri_golden.npz"))
    for tag in ("A", "B"):
        t = triangulate.multiTriangulator(g[tag + "_pts"].copy(), g[tag + "_intr"])
        xyz = t.triangulate()
        ref = g[tag + "_xyz"]
        assert np.array_equal(np.isnan(xyz), np.isnan(ref))
        ok = ~np.isnan(ref)
        assert np.abs(xyz[ok] - ref[ok]).max() <= 1e-8 * np.abs(ref[ok]).max()      # tolerance: 1e-8 relative
        assert np.abs(t.ext - g[tag + "_ext"]).max() < 1e-8
        assert list(t.ext[-1]) == [1, 0, 0, 0, 0, 0, 0]


@pytest.mark.parametrize("cfg", [(3, 500, 3), (5, 2000, 11), (2, 64, 4), (4, 9, 6)])
def test_pair_vs_oracle_and_opencv(cfg):
    import cv2
    m, n, seed = cfg
    s = scenes.make_ba_scene(m, n, m, seed=seed, radius=5.0)
    dense = _dense(s, m, n, seed, drop=0.0)
    intr = s["cams_true"][:, :10]
    for k in range(m - 1):
        args = (dense[:, 2 * k:2 * k + 2], dense[:, -2:], intr[k, 0], intr[-1, 0], intr[k, 1:3], intr[-1, 1:3], intr[k, -5:], intr[-1, -5:])
        T = triangulate.Triangulator(*args); xyz = T.triangulate()
        O = tri_port.Triangulator(*args); ref = O.triangulate()
        F = cv2.findFundamentalMat(O.p2, O.p1, method=cv2.FM_8POINT)[0]
        assert np.abs(T.F - F).max() <= 1e-8 * np.abs(F).max()
        assert np.abs(T.R - O.R).max() < 1e-8 and np.abs(T.t - O.t).max() < 1e-8
        assert np.abs(xyz - ref).max() <= 1e-8 * np.abs(ref).max()
        assert np.abs(T.getQuaternion() - O.getQuaternion()).max() < 1e-8


def test_too_few_points_is_an_error():
    s = scenes.make_ba_scene(2, 7, 2, seed=1, radius=5.0)
    dense = _dense(s, 2, 7, 1, drop=0.0)
    intr = s["cams_true"][:, :10]
    T = triangulate.Triangulator(dense[:, :2], dense[:, 2:], intr[0, 0], intr[1, 0], intr[0, 1:3], intr[1, 1:3], intr[0, -5:], intr[1, -5:])
    with pytest.raises(Exception):
        T.triangulate()


def test_large_pair_properties():
    """1M points seen by two cameras: the recovered pose reproduces the true relative pose (rotation exactly, translation up
    to the unknown scale) and the cloud is the true one in the reference camera's frame up to that scale."""
    m, n = 2, 1_000_000
    s = scenes.make_ba_scene(m, n, m, seed=5, radius=5.0)
    dense = _dense(s, m, n, 5, drop=0.0)
    intr = s["cams_true"][:, :10]
    T = triangulate.Triangulator(dense[:, :2], dense[:, 2:], intr[0, 0], intr[1, 0], intr[0, 1:3], intr[1, 1:3], intr[0, -5:], intr[1, -5:])
    xyz = T.triangulate()
    assert np.isfinite(xyz).all()
    assert abs(np.linalg.det(T.R) - 1.0) < 1e-6 and abs(np.linalg.norm(T.t) - 1.0) < 1e-9
    # a 2000-point subsample gives (nearly) the same pose and, for those points, the same cloud up to scale
    sub = np.random.default_rng(0).choice(n, 2000, replace=False)
    O = tri_port.Triangulator(dense[sub, :2], dense[sub, 2:], intr[0, 0], intr[1, 0], intr[0, 1:3], intr[1, 1:3], intr[0, -5:], intr[1, -5:])
    ref = O.triangulate()
    assert np.abs(T.R - O.R).max() < 5e-3
    sc = np.linalg.norm(xyz[sub], axis=1).mean() / np.linalg.norm(ref, axis=1).mean()
    assert np.median(np.linalg.norm(xyz[sub] - sc * ref, axis=1)) < 0.02 * np.linalg.norm(xyz[sub], axis=1).mean()
    # determinism
    assert np.array_equal(xyz, triangulate.Triangulator(dense[:, :2], dense[:, 2:], intr[0, 0], intr[1, 0], intr[0, 1:3], intr[1, 1:3],
                                                        intr[0, -5:], intr[1, -5:]).triangulate())
