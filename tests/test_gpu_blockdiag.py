"""GPU parity tests of the motion-only / structure-only LM: argus_ba_mot_kdrts / argus_ba_str_kdrts through
the C ABI against the golden runs of the reference's sba_mot_levmar_x / sba_str_levmar_x, against the CPU oracle on
seeded scenes, and through size-independent properties at BASELINE sizes."""
import numpy as np
import pytest

from argus_gui_b200 import scenes, ba
from oracle import ba_port

pytestmark = pytest.mark.gpu
M = 4
MOT_CASES = {"a": (0, 0, 0), "b": (2, 3, 1), "c": (5, 4, 0)}


def _info_match(info, ref, tol=1e-9):
    assert list(info[5:]) == list(ref[5:]), (info, ref)             # iterations, stop reason, nfev, njev, nlss: exact
    assert abs(info[0] - ref[0]) <= 1e-12 * ref[0]
    assert abs(info[1] - ref[1]) <= tol * ref[1]                    # tolerance: 1e-9 relative on the cost
    assert abs(info[4] - ref[4]) <= 1e-6 * abs(ref[4])


@pytest.mark.parametrize("tag", sorted(MOT_CASES))
@pytest.mark.parametrize("it", [1, 3, 10])
def test_mot_golden(golden_bd, tag, it):
    g = golden_bd
    nk, nd, mcon = MOT_CASES[tag]
    p0 = g["M1_p0"]
    cams, info, ret = ba.solve_mot(g["M1_vmask"], p0[:16 * M], p0[16 * M:], g["M1_rot0"], g["M1_x"], nk, nd, itmax=it, mcon=mcon)
    ref_c, ref_i = g["M1%s_cams_it%d" % (tag, it)], g["M1%s_info_it%d" % (tag, it)]
    if int(ref_i[6]) != 3:
        # the reference converged before itmax (case c: stop 2 after 8 iterations).  Once the cost stagnates, the stop-4 test
        # (sba_levmar.c:1903, eps4 = 0) compares (sqrt(e) - sqrt(e_new))^2 with 0 and fires on rounding noise, so the exact
        # iteration of the stop is not reproducible across implementations; the converged answer is.
        assert int(info[6]) in (2, 4) and abs(info[1] - ref_i[1]) <= 1e-8 * ref_i[1]
        assert np.abs(cams - ref_c).max() <= 1e-6 * np.abs(ref_c).max()
        return
    assert ret == int(ref_i[5])
    _info_match(info, ref_i)
    assert np.abs(cams - ref_c).max() <= 1e-8 * np.abs(ref_c).max()
    if mcon:
        assert np.array_equal(cams[:16 * mcon], p0[:16 * mcon])


@pytest.mark.parametrize("it", [2, 5])
def test_mot_rejected_steps_golden(golden_bd, it):
    g = golden_bd
    p0 = g["M2_p0"]
    cams, info, ret = ba.solve_mot(g["M2_vmask"], p0[:16 * M], p0[16 * M:], g["M2_rot0"], g["M2_x"], 0, 0, itmax=it, opts=g["M2_opts"])
    _info_match(info, g["M2_info_it%d" % it], tol=1e-8)
    assert np.abs(cams - g["M2_cams_it%d" % it]).max() <= 1e-7 * np.abs(g["M2_cams_it%d" % it]).max()


@pytest.mark.parametrize("tag,ncon", [("a", 0), ("b", 17)])
@pytest.mark.parametrize("it", [1, 2, 3])
def test_str_golden(golden_bd, tag, ncon, it):
    g = golden_bd
    p0 = g["M1_p0"]
    ref_p, ref_i = g["S1%s_pts_it%d" % (tag, it)], g["S1%s_info_it%d" % (tag, it)]
    # both equivalent camera forms: (rot0, zero local rotation) and the reference's (identity, full quaternion vector)
    for cams, rot0 in ((p0[:16 * M], g["M1_rot0"]), (g["S1_cams_full"], np.tile([1.0, 0.0, 0.0, 0.0], M))):
        pts, info, ret = ba.solve_str(g["M1_vmask"], p0[16 * M:], cams, rot0, g["M1_x"], itmax=it, ncon=ncon)
        _info_match(info, ref_i, tol=1e-10)
        assert np.abs(pts - ref_p).max() <= 1e-9 * np.abs(ref_p).max()
        if ncon:
            assert np.array_equal(pts[:3 * ncon], p0[16 * M:16 * M + 3 * ncon])


@pytest.mark.parametrize("it", [2, 4])
def test_str_rejected_steps_golden(golden_bd, it):
    g = golden_bd
    p0 = g["M2_p0"]
    pts, info, ret = ba.solve_str(g["M2_vmask"], p0[16 * M:], p0[:16 * M], g["M2_rot0"], g["M2_x"], itmax=it, opts=g["M2_opts"])
    _info_match(info, g["S2_info_it%d" % it], tol=1e-8)
    assert np.abs(pts - g["S2_pts_it%d" % it]).max() <= 1e-7 * np.abs(g["S2_pts_it%d" % it]).max()


@pytest.mark.parametrize("cfg", [(3, 400, 3, 5, 4, 0), (8, 3000, 5, 0, 0, 2), (17, 1500, 9, 2, 0, 0), (33, 900, 20, 0, 3, 1)])
def test_mot_vs_oracle(cfg):
    m, n, views, nk, nd, mcon = cfg
    s = scenes.make_ba_scene(m, n, views, seed=5 * m + n)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    for compat in (True, False):
        pp, ip, _ = ba_port.port_mot(p0, s["x"], s["vmask"], rot0, nk, nd, itmax=6, mcon=mcon, compat=compat)
        cams, info, ret = ba.solve_mot(s["vmask"], p0[:16 * m], p0[16 * m:], rot0, s["x"], nk, nd, itmax=6, mcon=mcon, compat=compat)
        _info_match(info, ip)
        assert np.abs(cams - pp[:16 * m]).max() <= 1e-8 * np.abs(pp[:16 * m]).max()


@pytest.mark.parametrize("cfg", [(3, 400, 3, 0), (8, 3000, 5, 40), (33, 900, 20, 0)])
def test_str_vs_oracle(cfg):
    m, n, views, ncon = cfg
    s = scenes.make_ba_scene(m, n, views, seed=7 * m + n)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    pp, ip, _ = ba_port.port_str(p0, s["x"], s["vmask"], rot0, itmax=3, ncon=ncon)
    pts, info, ret = ba.solve_str(s["vmask"], p0[16 * m:], p0[:16 * m], rot0, s["x"], itmax=3, ncon=ncon)
    _info_match(info, ip, tol=1e-10)
    assert np.abs(pts - pp[16 * m:]).max() <= 1e-9 * np.abs(pp[16 * m:]).max()


def test_resident_interface_and_errors():
    s = scenes.make_ba_scene(4, 500, 3, seed=9)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    B = ba.BundleAdjuster(mode="str").upload(s["vmask"], p0, rot0, s["x"])
    ret, info = B.solve(itmax=3)
    p1 = B.download()
    pts, info2, _ = ba.solve_str(s["vmask"], p0[64:], p0[:64], rot0, s["x"], itmax=3)
    assert np.array_equal(p1[64:], pts) and np.array_equal(p1[:64], p0[:64]) and list(info) == list(info2)
    B.reset(); ret2, info3 = B.solve(itmax=3)
    assert np.array_equal(B.download(), p1)                          # deterministic, and reset restores the start
    B.close()
    # fewer measurements than unknowns -> SBA_ERROR (sba_levmar.c:1586-1589), as in the reference
    vm = np.ones((1, 1), dtype=np.uint8)
    cam = np.zeros(16); cam[0] = 1000.0; cam[3] = 1.0; cam[15] = 5.0
    _, _, ret = ba.solve_mot(vm, cam, np.zeros(3), [1.0, 0, 0, 0], np.zeros(2))
    assert ret == -1


def test_full_size_properties():
    """BASELINE config 3 size (8 cameras, 200k points, 1.2M observations).  Structure-only with the TRUE cameras fixed
    recovers the true points to the noise level; motion-only with the TRUE points fixed brings the perturbed cameras back
    to the noise floor; both are deterministic run to run."""
    m, n, views = 8, 200_000, 6
    s = scenes.make_ba_scene(m, n, views, seed=3)
    pc, rot0 = scenes.pack_ba_problem(s["cams_true"], s["pts0"])
    pts, info, ret = ba.solve_str(s["vmask"], pc[16 * m:], pc[:16 * m], rot0, s["x"], itmax=20)
    assert ret > 0 and info[1] < info[0] and int(info[6]) in (1, 2, 4)
    err = np.linalg.norm(pts.reshape(n, 3) - s["pts_true"], axis=1)
    assert np.median(err) < 0.01 and info[1] / (2 * s["nobs"]) < 0.5       # 0.5 px noise -> ~0.25 px^2 per residual
    pts2, info2, _ = ba.solve_str(s["vmask"], pc[16 * m:], pc[:16 * m], rot0, s["x"], itmax=20)
    assert np.array_equal(pts, pts2) and list(info) == list(info2)
    pm, rot0m = scenes.pack_ba_problem(s["cams0"], s["pts_true"])
    cams, infom, retm = ba.solve_mot(s["vmask"], pm[:16 * m], pm[16 * m:], rot0m, s["x"], 0, 0, itmax=40)
    assert retm > 0 and infom[1] / (2 * s["nobs"]) < 0.5 and int(infom[6]) in (1, 2, 3, 4)
    cams2, infom2, _ = ba.solve_mot(s["vmask"], pm[:16 * m], pm[16 * m:], rot0m, s["x"], 0, 0, itmax=40)
    assert np.array_equal(cams, cams2) and list(infom) == list(infom2)


def test_sba_module_mot_and_struct_modes(golden_bd):
    """The python-sba style front end: options.motstruct = OPTS_MOT / OPTS_STRUCT reach the same drivers
    (the reference's demo switches the same way, sba-1.6p/demo/eucsbademo.c:1560-1602)."""
    from argus_gui_b200 import sba
    g = golden_bd
    p0, rot0 = g["M1_p0"], g["M1_rot0"].reshape(M, 4)
    cams17 = sba.Cameras.unpack(p0[:16 * M].reshape(M, 16), rot0)
    cameras = sba.Cameras.fromDylan(cams17)
    points = sba.Points(p0[16 * M:].reshape(-1, 3), g["M1_vmask"], g["M1_x"])
    opt = sba.Options.fromInput(cameras, points)
    opt.camera = sba.OPTS_CAMS; opt.nccalib = sba.OPTS_FIX2_INTR; opt.ncdist = sba.OPTS_FIX3_DIST
    opt.motstruct = sba.OPTS_MOT; opt.mcon = 1; opt.maxIter = 10
    nc, npts, info = sba.SparseBundleAdjust(cameras, points, opt)
    ref_i = g["M1b_info_it10"]
    assert info.iterations == int(ref_i[5]) and info.linSystems == int(ref_i[9]) and abs(info.finalError - ref_i[1]) <= 1e-9 * ref_i[1]
    assert np.array_equal(npts.B, points.B) and np.allclose(nc.arr[0], cameras.arr[0], rtol=1e-14, atol=1e-15)   # points and camera 0 fixed
    a_ref = g["M1b_cams_it10"].reshape(M, 16)
    assert np.abs(nc.arr[:, :10] - a_ref[:, :10]).max() <= 1e-8 * np.abs(a_ref[:, :10]).max()
    assert np.abs(nc.arr[:, 14:] - a_ref[:, 13:]).max() <= 1e-8
    opt.motstruct = sba.OPTS_STRUCT; opt.mcon = 0; opt.ncon = 17; opt.maxIter = 3
    nc, npts, info = sba.SparseBundleAdjust(cameras, points, opt)
    ref_i = g["S1b_info_it3"]
    assert info.iterations == 3 and info.linSystems == int(ref_i[9]) and abs(info.finalError - ref_i[1]) <= 1e-10 * ref_i[1]
    assert np.allclose(nc.arr, cameras.arr, rtol=1e-14, atol=1e-15)
    assert np.abs(npts.B.ravel() - g["S1b_pts_it3"]).max() <= 1e-9 * np.abs(g["S1b_pts_it3"]).max()
