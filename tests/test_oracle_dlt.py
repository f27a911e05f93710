"""CPU: the numpy DLT oracle (oracle/dlt_port.py) against golden vectors produced by the reference's own
argus_gui/tools.py (tests/golden/make_dlt_golden.py)."""
import numpy as np
from oracle import dlt_port


def test_undistort_bit_exact_float32(golden_dlt):
    g = golden_dlt
    out = dlt_port.undistort_pts_f32(g["U_px"], g["U_prof"])
    assert out.dtype == np.float32
    assert np.array_equal(out, g["U_out"])


def _check_case(g, c, m, k):
    pts, prof, dlt = g[c + "_pts"], g[c + "_prof"], g[c + "_dlt"]
    xyz = np.hstack([dlt_port.uv_to_xyz(pts[:, t * 2 * m:(t + 1) * 2 * m], prof, dlt) for t in range(k)])
    ref = g[c + "_xyz"]
    assert np.array_equal(np.isnan(xyz), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.abs(xyz[ok] - ref[ok]).max() <= 1e-12 * np.abs(ref[ok]).max()
    err = dlt_port.get_repo_errors(ref, pts, prof, dlt)
    rerr = g[c + "_err"]
    assert err.shape == rerr.shape
    assert np.array_equal(np.isnan(err), np.isnan(rerr))
    ok = ~np.isnan(rerr)
    assert np.abs(err[ok] - rerr[ok]).max() <= 1e-12 * np.abs(rerr[ok]).max()


def test_uv_to_xyz_and_errors_case_a(golden_dlt):
    _check_case(golden_dlt, "A", 4, 3)


def test_uv_to_xyz_and_errors_case_b(golden_dlt):
    _check_case(golden_dlt, "B", 6, 2)


def test_quirk_differs_from_textbook(golden_dlt):
    """The reference's (n+1)-row system is NOT the textbook 2n-row one."""
    g = golden_dlt
    pts, prof, dlt = g["A_pts"][:50, :8], g["A_prof"], g["A_dlt"]
    a = dlt_port.uv_to_xyz(pts, prof, dlt)
    b = dlt_port.uv_to_xyz(pts, prof, dlt, textbook=True)
    ok = ~np.isnan(a).any(axis=1) & ~np.isnan(b).any(axis=1)
    assert np.abs(a[ok] - b[ok]).max() > 1e-6


def test_edge_rows(golden_dlt):
    g = golden_dlt
    xyz = g["A_xyz"]
    assert np.isnan(xyz[0, :3]).all()        # nothing visible
    assert np.isnan(xyz[1, :3]).all()        # a single camera
    assert not np.isnan(xyz[2, :3]).any()    # exactly two cameras


def test_bootstrap_sd_golden():
    """oracle restatement of the bootstrap loop (tools.py:513-530) against the reference run with the same random numbers"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bootstrap_golden.npz"))
    ci = 1.96 * dlt_port.bootstrap_sd(g["pts"], g["rmses"], g["prof"], g["dlt"], g["noise"])
    assert np.array_equal(np.isnan(ci), np.isnan(g["ci"]))
    ok = ~np.isnan(g["ci"])
    assert np.abs(ci[ok] - g["ci"][ok]).max() <= 1e-12 * np.abs(g["ci"][ok]).max()


def test_oracle_solve_dlt_golden():
    """numpy restatement of the per-camera DLT fit against golden output of the reference's OutlierWindow.getCoefficients
    (tests/golden/make_fit_golden.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fit_golden.npz"))
    for tag in ("A", "B", "C"):
        for j in range(3):
            k = "%s%d" % (tag, j)
            L, err, idx = dlt_port.solve_dlt(g[k + "_xyz"], g[k + "_uv"], both_coordinates=False)
            assert np.abs(L - g[k + "_L"]).max() <= 1e-12 * np.abs(g[k + "_L"]).max()
            assert abs(err.mean() - float(g[k + "_rmse"])) <= 1e-9 * float(g[k + "_rmse"]) + 1e-12
            if tag != "B":
                assert list(idx[err >= 3 * err.std() + err.mean()]) == list(g[k + "_flagged"])
