"""In-solver wand-length constraint (argus_ba_set_wand, csrc/ba_wand.cuh).  EXTENSION - PARITY UNPINNED: the reference applies
the wand length after the adjustment only (/root/reference/argus_gui/sbaDriver.py:703-713), so these tests check the CUDA path
against a dense numpy statement of the same mathematics (oracle/wand_numpy.py), against finite differences (the style of the
reference's Jacobian checker, sba-1.6p/sba_chkjac.c:91-211), and through properties: switched off it is the unconstrained
solver bit for bit, switched on the wand lengths are pulled to the known length."""
import numpy as np
import pytest

from argus_gui_b200 import scenes, ba
from oracle import ba_port, wand_numpy
from helpers import rel

pytestmark = pytest.mark.gpu


def _wand_lengths(p, m, s):
    pts = p[16 * m:].reshape(-1, 3)
    return np.linalg.norm(pts[s["ia"]] - pts[s["ib"]], axis=1)


def test_switched_off_is_bit_identical():
    m = 4
    s = scenes.make_ba_wand_scene(m, 150, 40, 3, seed=5)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    out = []
    for how in ("never set", "weight 0", "no pairs"):
        B = ba.BundleAdjuster().upload(s["vmask"], p0, rot0, s["x"], 2, 3)
        if how == "weight 0":
            B.set_wand(s["ia"], s["ib"], s["wand"], 0.0)
        elif how == "no pairs":
            B.set_wand([], [], s["wand"], 100.0)
        ret, info = B.solve(itmax=8)
        out.append((ret, info.copy(), B.download().copy())); B.close()
    for ret, info, p in out[1:]:
        assert ret == out[0][0] and info.tobytes() == out[0][1].tobytes() and p.tobytes() == out[0][2].tobytes()


def test_finite_differences_of_the_wand_rows():
    """CHKDER-style: the analytic rows of the dense model against central differences of the wand residuals."""
    rng = np.random.default_rng(3)
    pts = rng.normal(size=(12, 3)); ia = np.arange(6); ib = np.arange(6, 12); w, L = 37.0, 0.5
    g = wand_numpy.wand_jacobian_rows(pts, ia, ib, w)
    for p in range(6):
        for t in range(3):
            h = 1e-6
            q = pts.copy(); q[ia[p], t] += h
            rp = wand_numpy.wand_residuals(q, ia, ib, L, w)[p]
            q[ia[p], t] -= 2 * h
            rm = wand_numpy.wand_residuals(q, ia, ib, L, w)[p]
            assert abs(-(rp - rm) / (2 * h) - g[p, t]) < 1e-6 * w            # J = d prediction / d p = -d residual / d p


@pytest.mark.parametrize("cfg", [(3, 30, 12, 3, 0, 0, 0, 0), (4, 40, 25, 3, 2, 3, 1, 0), (5, 25, 30, 4, 0, 0, 1, 7), (3, 20, 0, 3, 5, 4, 0, 0)])
def test_one_step_against_the_dense_model(cfg):
    """Cost, ||J^T e||_inf, max diagonal, the step dp (cameras and points, paired and unpaired, fixed ends), dL and the cost at
    the trial point: the Schur-complement + Sherman-Morrison path on the GPU against dense (J^T J + mu I) dp = J^T e."""
    m, npairs, nextra, views, nk, nd, mcon, ncon = cfg
    s = scenes.make_ba_wand_scene(m, npairs, nextra, views, seed=11 * m + npairs)
    n = 2 * npairs + nextra
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    # ncon fixes the FIRST points: ends 1 of the first pairs - pairs with one fixed end are part of the case
    w, mu = 120.0, 0.37
    A, Bj = ba_port.port_jacobian(p0, s["vmask"], rot0, nk, nd)
    hx = ba_port.port_project(p0, s["vmask"], rot0)
    pi, ci = np.nonzero(s["vmask"])
    pts0 = p0[16 * m:].reshape(n, 3)
    ds = wand_numpy.dense_step(A, Bj, s["x"] - hx, pi, ci, n, m, mu, pts0, s["ia"], s["ib"], s["wand"], w, mcon=mcon, ncon=ncon)
    B = ba.BundleAdjuster().upload(s["vmask"], p0, rot0, s["x"], nk, nd, ncon=ncon, mcon=mcon)
    B.set_wand(s["ia"], s["ib"], s["wand"], w)
    g = B.normal_equations(mu, mcon=mcon)
    B.close()
    sc = g["scal"]
    assert abs(sc[0] - ds["cost"]) <= 1e-12 * ds["cost"]
    assert abs(sc[1] - np.abs(ds["JTe"]).max()) <= 1e-10 * np.abs(ds["JTe"]).max()
    assert abs(sc[3] - ds["diag"].max()) <= 1e-11 * ds["diag"].max()
    assert sc[7] == 1.0
    dpg = g["dp"]
    cam_scale = np.abs(ds["dp"][:16 * m]).max(); pt_scale = np.abs(ds["dp"][16 * m:]).max()
    assert np.abs(dpg[:16 * m] - ds["dp"][:16 * m]).max() <= 1e-7 * cam_scale      # dense solve of an ill-conditioned system: 1e-7
    assert np.abs(dpg[16 * m:] - ds["dp"][16 * m:]).max() <= 1e-7 * pt_scale
    assert abs(sc[4] - ds["dp"] @ ds["dp"]) <= 1e-7 * (ds["dp"] @ ds["dp"])
    assert abs(sc[5] - ds["dL"]) <= 1e-7 * abs(ds["dL"])
    # cost at the trial point, recomputed on the CPU at the GPU's own step
    pn = p0 + dpg
    en = s["x"] - ba_port.port_project(pn, s["vmask"], rot0)
    rn = wand_numpy.wand_residuals(pn[16 * m:].reshape(n, 3), s["ia"], s["ib"], s["wand"], w)
    ref = float((en * en).sum() + rn @ rn)
    assert abs(sc[6] - ref) <= 1e-11 * ref


def test_wand_lengths_are_pulled_to_the_known_length():
    """C1-shaped (3 cameras, intrinsics fixed) with coarse pixel noise: with the constraint the spread of the reconstructed wand
    lengths (the reference's wand score, sbaDriver.py:779-781) shrinks and their mean sits at the known length, the cost is
    monotone and the solver stops by itself."""
    m = 3
    s = scenes.make_ba_wand_scene(m, 400, 50, 3, seed=2, noise=1.5, dist=(-0.05, 0, 0, 0, 0))
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    res = {}
    for w in (0.0, 2000.0):
        B = ba.BundleAdjuster().upload(s["vmask"], p0, rot0, s["x"], 5, 4, mcon=1)
        B.set_wand(s["ia"], s["ib"], s["wand"], w)
        ret, info = B.solve(itmax=100)
        res[w] = (ret, info, _wand_lengths(B.download(), m, s)); B.close()
        assert ret > 0 and info[1] < info[0] and int(info[6]) in (2, 4, 6)
    d0, d1 = res[0.0][2], res[2000.0][2]
    assert d1.std() < 0.2 * d0.std()
    assert abs(d1.mean() - s["wand"]) < 1e-3 * s["wand"]
    # the reprojection part of the cost can only get worse under the constraint, and not by much
    assert res[2000.0][1][1] >= res[0.0][1][1] * (1 - 1e-9)


def test_rejected_steps_and_graph_replay_with_the_constraint():
    """A far start with a tiny initial damping (steps get rejected) through the CUDA-graph path twice: same answer both times."""
    m = 4
    s = scenes.make_ba_wand_scene(m, 200, 30, 4, seed=9, pert_scale=30.0)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    B = ba.BundleAdjuster().upload(s["vmask"], p0, rot0, s["x"], 2, 3, mcon=1)
    B.set_wand(s["ia"], s["ib"], s["wand"], 500.0)
    runs = []
    for _ in range(2):
        B.reset()
        ret, info = B.solve(itmax=12, opts=[1e-8, 1e-12, 1e-12, 1e-12, 0.0])
        runs.append((ret, info.copy(), B.download().copy()))
    B.close()
    assert runs[0][1][9] >= runs[0][1][5] and runs[0][1][1] < runs[0][1][0]
    assert runs[0][0] == runs[1][0] and runs[0][1].tobytes() == runs[1][1].tobytes() and runs[0][2].tobytes() == runs[1][2].tobytes()


def test_bad_pair_lists_are_refused():
    m = 3
    s = scenes.make_ba_wand_scene(m, 20, 5, 3, seed=1)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    B = ba.BundleAdjuster().upload(s["vmask"], p0, rot0, s["x"], 5, 4)
    for ia, ib in (([0, 0], [1, 2]), ([0], [0]), ([0], [45]), ([-1], [3])):
        with pytest.raises(ValueError):
            B.set_wand(ia, ib, 0.5, 10.0)
    B.close()
