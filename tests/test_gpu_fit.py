"""GPU parity tests of the per-camera DLT fit: argus_dlt_fit through the C ABI / tools.solve_dlt against
golden output of the reference's own getCoefficients, against the numpy restatement (lstsq) on seeded data, and at a
large size through properties."""
import os
import numpy as np
import pytest

from argus_gui_b200 import scenes, tools, dlt
from oracle import dlt_port

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fit_golden():
    """Golden output of the reference's OutlierWindow.getCoefficients (its tools.solve_dlt twin does not run under numpy 2):
    coefficients, mean reprojection distance and the indices its 3-sigma rule flags."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "fit_golden.npz"))
    for tag in ("A", "B", "C"):
        for j in range(3):
            k = "%s%d" % (tag, j)
            xyz, uv = g[k + "_xyz"], g[k + "_uv"]
            L, err, (n, esum, mean, std) = dlt.fit(xyz, uv)
            ref = g[k + "_L"]
            assert np.abs(L - ref).max() <= 1e-8 * np.abs(ref).max()               # tolerance: 1e-8 relative (QR vs LAPACK SVD)
            rm = float(g[k + "_rmse"])
            assert abs(esum / n - rm) <= 1e-8 * rm + 1e-9                           # pixels; case B is noise-free (distances ~1e-11)
            if tag != "B":                                                          # B is noise-free: distances are rounding noise
                flagged = np.nonzero(err >= 3 * std + mean)[0]
                assert list(flagged) == list(g[k + "_flagged"])
            # tools.solve_dlt on rows without NaN is the same fit
            ok = ~np.isnan(uv).any(axis=1)
            L2, rmse2 = tools.solve_dlt(xyz[ok].copy(), uv[ok].copy())
            assert L2.shape == (11, 1) and np.abs(L2[:, 0] - ref).max() <= 1e-8 * np.abs(ref).max() and abs(rmse2 - rm) <= 1e-8 * rm + 1e-9


@pytest.mark.parametrize("n,seed,both", [(7, 1, True), (200, 2, False), (20_000, 3, True), (300_001, 4, False)])
def test_fit_vs_numpy_lstsq(n, seed, both):
    s = scenes.make_dlt_scene(n, m=2, ntracks=1, seed=seed, nan_frac=0.0)
    rng = np.random.default_rng(seed)
    xyz = s["xyz_true"].reshape(n, 3)
    L0 = s["dlt"][1]
    den = xyz @ L0[8:11] + 1.0
    uv = np.stack([(xyz @ L0[0:3] + L0[3]) / den, (xyz @ L0[4:7] + L0[7]) / den], axis=1) + rng.normal(0, 0.7, (n, 2))
    if n > 50:
        uv[rng.random(n) < 0.2, 0] = np.nan
        uv[rng.random(n) < 0.05, 1] = np.nan
        if not both:
            uv[np.isnan(uv[:, 1]), 0] = np.nan     # the u-only rule must never see a lone NaN in v (it would poison the reference too)
    L, err, (nu, esum, mean, std) = dlt.fit(xyz, uv, both_coordinates=both)
    Lr, er, idx = dlt_port.solve_dlt(xyz, uv, both_coordinates=both)
    assert nu == len(idx)
    assert np.abs(L - Lr).max() <= 1e-8 * np.abs(Lr).max()
    assert np.array_equal(np.nonzero(~np.isnan(err))[0], idx)
    assert np.abs(err[idx] - er).max() <= 1e-7 * max(er.max(), 1.0)
    assert abs(esum - er.sum()) <= 1e-9 * er.sum() and abs(mean - er.mean()) <= 1e-9 * er.mean() and abs(std - er.std()) <= 1e-7 * er.std()


def test_rank_deficient_is_an_error():
    xyz = np.zeros((20, 3)); xyz[:, 0] = np.arange(20)          # collinear points
    uv = np.stack([np.arange(20) * 3.0, np.arange(20) * 2.0], axis=1)
    with pytest.raises(ArithmeticError):
        dlt.fit(xyz, uv)


def test_exact_data_recovers_coefficients_and_is_deterministic():
    n = 2_000_000
    s = scenes.make_dlt_scene(n, m=2, ntracks=1, seed=9, nan_frac=0.0)
    xyz = s["xyz_true"].reshape(n, 3)
    L0 = s["dlt"][0]
    den = xyz @ L0[8:11] + 1.0
    uv = np.stack([(xyz @ L0[0:3] + L0[3]) / den, (xyz @ L0[4:7] + L0[7]) / den], axis=1)
    L, err, st = dlt.fit(xyz, uv)
    assert np.abs(L - L0).max() <= 1e-9 * np.abs(L0).max() and np.nanmax(err) < 1e-8
    L2, err2, st2 = dlt.fit(xyz, uv)
    assert np.array_equal(L, L2) and np.array_equal(err, err2) and st == st2
