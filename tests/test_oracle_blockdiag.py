"""CPU: the restatement of the motion-only / structure-only LM (oracle/ba_port.c orc_blockdiag) against the golden runs of
the reference's own sba_mot_levmar_x / sba_str_levmar_x (tests/golden/make_blockdiag_golden.py), and - where oracle/_ref
is present - against the reference run live."""
import numpy as np
import pytest

from oracle import ba_port, sba_ref

M = 4
MOT_CASES = {"a": (0, 0, 0), "b": (2, 3, 1), "c": (5, 4, 0)}


def _info_match(info, ref, tol=1e-9):
    assert list(info[5:]) == list(ref[5:]), (info, ref)             # iterations, stop reason, nfev, njev, nlss: exact
    assert abs(info[0] - ref[0]) <= 1e-12 * ref[0]
    assert abs(info[1] - ref[1]) <= tol * ref[1]
    assert abs(info[4] - ref[4]) <= 1e-6 * abs(ref[4])


@pytest.mark.parametrize("tag", sorted(MOT_CASES))
@pytest.mark.parametrize("it", [1, 3, 10])
def test_port_mot_golden(golden_bd, tag, it):
    g = golden_bd
    nk, nd, mcon = MOT_CASES[tag]
    p, info, ret = ba_port.port_mot(g["M1_p0"], g["M1_x"], g["M1_vmask"], g["M1_rot0"], nk, nd, itmax=it, mcon=mcon)
    ref_c, ref_i = g["M1%s_cams_it%d" % (tag, it)], g["M1%s_info_it%d" % (tag, it)]
    _info_match(info, ref_i)
    assert np.abs(p[:16 * M] - ref_c).max() <= 1e-8 * np.abs(ref_c).max()
    assert np.array_equal(p[16 * M:], g["M1_p0"][16 * M:])          # the points do not move
    if mcon:
        assert np.array_equal(p[:16 * mcon], g["M1_p0"][:16 * mcon])


@pytest.mark.parametrize("it", [2, 5])
def test_port_mot_rejected_steps_golden(golden_bd, it):
    g = golden_bd
    p, info, ret = ba_port.port_mot(g["M2_p0"], g["M2_x"], g["M2_vmask"], g["M2_rot0"], 0, 0, itmax=it, opts=g["M2_opts"])
    ref_i = g["M2_info_it%d" % it]
    assert ref_i[7] > ref_i[5] + 1                                   # the fixture does contain rejected steps
    _info_match(info, ref_i, tol=1e-8)
    assert np.abs(p[:16 * M] - g["M2_cams_it%d" % it]).max() <= 1e-7 * np.abs(g["M2_cams_it%d" % it]).max()


@pytest.mark.parametrize("tag,ncon", [("a", 0), ("b", 17)])
@pytest.mark.parametrize("it", [1, 2, 3])
def test_port_str_golden(golden_bd, tag, ncon, it):
    """The reference's structure-only callbacks take the full camera quaternion from the camera row; the port is run in
    both equivalent forms: (rot0, zero local rotation) and (identity rot0, full quaternion vector)."""
    g = golden_bd
    ref_p, ref_i = g["S1%s_pts_it%d" % (tag, it)], g["S1%s_info_it%d" % (tag, it)]
    pid = np.concatenate([g["S1_cams_full"], g["M1_p0"][16 * M:]])
    for p0, rot0 in ((g["M1_p0"], g["M1_rot0"]), (pid, np.tile([1.0, 0.0, 0.0, 0.0], M))):
        p, info, ret = ba_port.port_str(p0, g["M1_x"], g["M1_vmask"], rot0, itmax=it, ncon=ncon)
        _info_match(info, ref_i, tol=1e-10)
        assert np.abs(p[16 * M:] - ref_p).max() <= 1e-9 * np.abs(ref_p).max()
        assert np.array_equal(p[:16 * M], p0[:16 * M])               # the cameras do not move
        if ncon:
            assert np.array_equal(p[16 * M:16 * M + 3 * ncon], p0[16 * M:16 * M + 3 * ncon])


@pytest.mark.parametrize("it", [2, 4])
def test_port_str_rejected_steps_golden(golden_bd, it):
    g = golden_bd
    p, info, ret = ba_port.port_str(g["M2_p0"], g["M2_x"], g["M2_vmask"], g["M2_rot0"], itmax=it, opts=g["M2_opts"])
    _info_match(info, g["S2_info_it%d" % it], tol=1e-8)
    assert np.abs(p[16 * M:] - g["S2_pts_it%d" % it]).max() <= 1e-7 * np.abs(g["S2_pts_it%d" % it]).max()


def test_too_few_measurements_is_an_error():
    vm = np.ones((1, 1), dtype=np.uint8)
    p = np.zeros(19); p[0] = 1000.0; p[3] = 1.0; p[15] = 5.0
    _, _, ret = ba_port.port_mot(p, np.zeros(2), vm, [1.0, 0, 0, 0], 0, 0)          # 2 measurements < 16 unknowns
    assert ret == -1


@pytest.mark.skipif(not sba_ref.available(), reason="oracle/_ref not built")
def test_port_vs_reference_live():
    from argus_gui_b200 import scenes
    s = scenes.make_ba_scene(5, 250, 4, seed=77)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    m = 5
    cr, ir, _ = sba_ref.ref_mot(p0[:16 * m], p0[16 * m:], s["x"], s["vmask"], rot0, 4, 5, itmax=6, mcon=2)
    pp, ip, _ = ba_port.port_mot(p0, s["x"], s["vmask"], rot0, 4, 5, itmax=6, mcon=2)
    _info_match(ip, ir)
    assert np.abs(pp[:16 * m] - cr).max() <= 1e-8 * np.abs(cr).max()
