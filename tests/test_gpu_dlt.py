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
    # the two-call path (uv_to_xyz then get_repo_errors) over several transfer chunks gives the fused call's numbers
    sub = np.ascontiguousarray(pts[:120_000])
    x3 = dlt.triangulate(sub, d["prof"], d["dlt"])
    e3 = dlt.reproj_error(x3, sub, d["prof"], d["dlt"])
    assert np.array_equal(x3, xyz[:120_000], equal_nan=True) and np.array_equal(e3, err[:, :120_000], equal_nan=True)
    # preallocated outputs
    xo_, eo_ = np.empty((120_000, 3 * k)), np.empty((k, 120_000))
    dlt.reconstruct(sub, d["prof"], d["dlt"], out=(xo_, eo_))
    assert np.array_equal(xo_, x3, equal_nan=True) and np.array_equal(eo_, e3, equal_nan=True)
    # oracle on a sample of frames
    sidx = idx[:300]
    for t in (0, 7):
        xo = dlt_port.uv_to_xyz(pts[sidx][:, t * 2 * m:(t + 1) * 2 * m], d["prof"], d["dlt"])
        okk = ~np.isnan(xo)
        xs = xyz[sidx][:, 3 * t:3 * t + 3]
        assert np.array_equal(np.isnan(xs), np.isnan(xo)) and np.abs(xs[okk] - xo[okk]).max() <= 1e-9 * np.abs(xo[okk]).max()


def _boot_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bootstrap_golden.npz"))


def test_bootstrap_exact_noise_golden():
    """bootstrapXYZs (tools.py:448-544) with the reference's own random numbers: CIs, weights and tolerances."""
    from argus_gui_b200 import tools
    g = _boot_golden()
    ci, w, tol = tools.bootstrapXYZs(g["pts"].copy(), g["rmses"], list(g["prof"]), g["dlt"], bsIter=g["noise"].shape[1], subframeinterp=False,
                                     noise=g["noise"])
    assert np.array_equal(np.isnan(ci), np.isnan(g["ci"]))
    ok = ~np.isnan(g["ci"])
    assert np.abs(ci[ok] - g["ci"][ok]).max() <= 1e-8 * np.abs(g["ci"][ok]).max()      # standard deviations: 1e-8 relative
    okw = ~np.isnan(g["w"])
    assert np.array_equal(np.isnan(w), np.isnan(g["w"])) and np.abs(w[okw] - g["w"][okw]).max() <= 1e-7 * np.abs(g["w"][okw]).max()
    assert np.abs(tol - g["tol"]).max() <= 1e-7 * np.abs(g["tol"]).max()


def test_bootstrap_subframe_interpolation_golden():
    """The sub-frame offset search (tools.py:455-510) finds the same offset, rewrites pts the same way, same CIs."""
    from argus_gui_b200 import tools
    g = _boot_golden()
    pts = g["pts2_in"].copy()
    ci, w, tol = tools.bootstrapXYZs(pts, g["rm2"], list(g["prof"]), g["dlt"], bsIter=g["noise2"].shape[1], subframeinterp=True, noise=g["noise2"])
    assert np.abs(pts - g["pts2_out"]).max() < 1e-9                 # in-place rewrite of the shifted camera
    ok = ~np.isnan(g["ci2"])
    assert np.array_equal(np.isnan(ci), np.isnan(g["ci2"]))
    assert np.abs(ci[ok] - g["ci2"][ok]).max() <= 1e-6 * np.abs(g["ci2"][ok]).max()


def test_bootstrap_device_rng_statistics():
    """On-device Philox noise: the standard deviations agree statistically with the explicit-noise estimate."""
    g = _boot_golden()
    rng = np.random.default_rng(0)
    N, m2 = g["pts"].shape[0], g["pts"].shape[1] // 2
    bs = 4000
    a = dlt.bootstrap_sd(g["pts"], g["rmses"], g["prof"], g["dlt"], bs_iter=bs, seed=1234)
    b = dlt.bootstrap_sd(g["pts"], g["rmses"], g["prof"], g["dlt"], bs_iter=bs, seed=99)
    noise = rng.standard_normal((2, 400, N, 8))
    c = dlt.bootstrap_sd(g["pts"], g["rmses"], g["prof"], g["dlt"], bs_iter=400, noise=noise)
    ok = ~np.isnan(a) & ~np.isnan(c)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.median(np.abs(a[ok] / b[ok] - 1)) < 0.03              # two seeds, 4000 draws each: ~1.6 % standard error
    assert np.median(np.abs(a[ok] / c[ok] - 1)) < 0.08              # vs 400 explicit numpy draws: ~3.5 %
    assert not np.array_equal(a, b)


@pytest.mark.parametrize("m,k", [(2, 1), (3, 5), (8, 2), (12, 1), (16, 3), (17, 2), (40, 1), (64, 1)])
def test_camera_count_paths_vs_oracle(m, k):
    """Every kernel variant: fused <4>, <8>, <16> and the two-kernel path for more than 16 cameras, up to the 64-camera limit;
    the fused call equals the two separate calls bit for bit in all of them."""
    N = 120
    d = scenes.make_dlt_scene(N, m=m, ntracks=k, seed=100 + m, nan_frac=0.25)
    pts = d["pts"]
    xyz, err = dlt.reconstruct(pts, d["prof"], d["dlt"])
    x1 = dlt.triangulate(pts, d["prof"], d["dlt"]); e1 = dlt.reproj_error(x1, pts, d["prof"], d["dlt"])
    assert np.array_equal(xyz, x1, equal_nan=True) and np.array_equal(err, e1, equal_nan=True)
    for t in range(k):
        sl = pts[:, t * 2 * m:(t + 1) * 2 * m]
        xo = dlt_port.uv_to_xyz(sl, d["prof"], d["dlt"])
        assert np.array_equal(np.isnan(xo), np.isnan(xyz[:, 3 * t:3 * t + 3]))
        ok = ~np.isnan(xo)
        assert np.abs(xyz[:, 3 * t:3 * t + 3][ok] - xo[ok]).max() <= 1e-9 * np.abs(xo[ok]).max()
    eo = dlt_port.get_repo_errors(xyz, pts, d["prof"], d["dlt"])
    ok = ~np.isnan(eo)
    assert np.array_equal(np.isnan(eo), np.isnan(err)) and np.abs(err[ok] - eo[ok]).max() <= 1e-9 * np.abs(eo[ok]).max()
    if m == 64:
        with pytest.raises(ValueError):
            d65 = scenes.make_dlt_scene(4, m=65, ntracks=1, seed=1)
            dlt.triangulate(d65["pts"], d65["prof"], d65["dlt"])
