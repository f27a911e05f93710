"""GPU parity tests of the DLT path (uv_to_xyz, get_repo_errors, undistort_pts) through the C ABI:
golden vectors from the reference's own tools.py, the numpy oracle, OpenCV itself, and full-size (config 2) properties."""
import numpy as np
import pytest
from argus_gui_b200 import dlt, scenes
from oracle import dlt_port

pytestmark = pytest.mark.gpu


def test_d0_undistort_bit_exact(golden_dlt):
    g = golden_dlt
    out = dlt.undistort(g["U_px"], g["U_prof"])
    assert out.dtype == np.float32 and np.array_equal(out, g["U_out"])       # bit-exact float32 (D0)


def test_d0_undistort_vs_opencv_live():
    import cv2
    rng = np.random.default_rng(9)
    px = rng.uniform(-300, 2300, size=(20000, 2))
    prof = np.array([1234.5, 950.0, 530.0, -0.21, 0.07, 0.002, -0.0015, -0.008])
    K = np.array([[prof[0], 0, prof[1]], [0, prof[0], prof[2]], [0, 0, 1.0]])
    src = np.zeros((1, px.shape[0], 2), dtype=np.float32); src[0] = px
    ref = cv2.undistortPoints(src, K, prof[-5:], P=K).reshape(-1, 2)
    assert np.array_equal(dlt.undistort(px, prof), ref)


@pytest.mark.parametrize("case,m,k", [("A", 4, 3), ("B", 6, 2)])
def test_d1_d2_golden(golden_dlt, case, m, k):
    g = golden_dlt
    pts, prof, L = g[case + "_pts"], g[case + "_prof"], g[case + "_dlt"]
    xyz = dlt.triangulate(pts, prof, L)
    ref = g[case + "_xyz"]
    assert np.array_equal(np.isnan(xyz), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.abs(xyz[ok] - ref[ok]).max() <= 1e-9 * np.abs(ref[ok]).max()          # 1e-9 (D1)
    err = dlt.reproj_error(ref, pts, prof, L)
    rerr = g[case + "_err"]
    assert err.shape == rerr.shape and np.array_equal(np.isnan(err), np.isnan(rerr))
    ok = ~np.isnan(rerr)
    assert np.abs(err[ok] - rerr[ok]).max() <= 1e-9 * np.abs(rerr[ok]).max()          # 1e-9 (D2)
    # the per-track call of the reference (k = 1 slices) gives the same numbers as the batched call
    for t in range(k):
        one = dlt.triangulate(np.ascontiguousarray(pts[:, t * 2 * m:(t + 1) * 2 * m]), prof, L)
        assert np.array_equal(one, xyz[:, 3 * t:3 * t + 3], equal_nan=True)
    x2, e2 = dlt.reconstruct(pts, prof, L)          # fused kernel: identical to the two separate calls, bit for bit
    assert np.array_equal(x2, xyz, equal_nan=True)
    assert np.array_equal(e2, dlt.reproj_error(xyz, pts, prof, L), equal_nan=True)


def test_textbook_flag_matches_oracle(golden_dlt):
    g = golden_dlt
    pts, prof, L = np.ascontiguousarray(g["A_pts"][:, :8]), g["A_prof"], g["A_dlt"]
    a = dlt.triangulate(pts, prof, L, textbook=True)
    b = dlt_port.uv_to_xyz(pts, prof, L, textbook=True)
    ok = ~np.isnan(b)
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.abs(a[ok] - b[ok]).max() <= 1e-9 * np.abs(b[ok]).max()


def test_empty_and_all_nan():
    d = scenes.make_dlt_scene(10, m=4, ntracks=2, seed=1)
    xyz = dlt.triangulate(np.zeros((0, 16)), d["prof"], d["dlt"])
    assert xyz.shape == (0, 6)
    pts = np.full((7, 16), np.nan)
    xyz, err = dlt.reconstruct(pts, d["prof"], d["dlt"])
    assert np.isnan(xyz).all() and np.isnan(err).all()
    with pytest.raises(ValueError):
        dlt.triangulate(np.zeros((5, 7)), d["prof"], d["dlt"])


def test_full_size_c2_properties():
    """BASELINE config 2: 1M frames x 4 cameras x 10 tracks.  A random sample of frames is checked against the oracle;
    the whole output is checked through size-independent properties."""
    N, m, k = 1_000_000, 4, 10
    d = scenes.make_dlt_scene(N, m=m, ntracks=k, seed=2)
    pts = d["pts"]
    xyz, err = dlt.reconstruct(pts, d["prof"], d["dlt"])
    assert xyz.shape == (N, 3 * k) and err.shape == (k, N)
    vis = (~np.isnan(pts.reshape(N, k, m, 2)).any(axis=3)).sum(axis=2)        # cameras per (frame, track)
    assert np.array_equal(np.isnan(xyz.reshape(N, k, 3)).any(axis=2), vis < 2)   # NaN exactly where < 2 cameras
    assert np.array_equal(np.isnan(err.T), vis < 2)
    # recovered points are the true ones up to the pixel noise (0.5 px at ~5 m with f = 1000 -> millimetres)
    ok = vis >= 3
    dist = np.linalg.norm(xyz.reshape(N, k, 3) - d["xyz_true"].reshape(N, k, 3), axis=2)
    assert np.median(dist[ok]) < 0.01 and dist[ok].max() < 0.5
    assert np.nanmedian(err) < 2.0
    # frames are independent: a shuffled batch gives the shuffled result, bit for bit
    idx = np.random.default_rng(1).permutation(N)[:50_000]
    x2, e2 = dlt.reconstruct(np.ascontiguousarray(pts[idx]), d["prof"], d["dlt"])
    assert np.array_equal(x2, xyz[idx], equal_nan=True) and np.array_equal(e2, err[:, idx], equal_nan=True)
    # oracle on a sample of frames
    sidx = idx[:300]
    for t in (0, 7):
        xo = dlt_port.uv_to_xyz(pts[sidx][:, t * 2 * m:(t + 1) * 2 * m], d["prof"], d["dlt"])
        okk = ~np.isnan(xo)
        xs = xyz[sidx][:, 3 * t:3 * t + 3]
        assert np.array_equal(np.isnan(xs), np.isnan(xo)) and np.abs(xs[okk] - xo[okk]).max() <= 1e-9 * np.abs(xo[okk]).max()
