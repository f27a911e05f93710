"""CPU: the C restatement of the reference BA (oracle/ba_port.c) against
 (a) golden vectors produced by the reference's own compiled code (tests/golden/ba_golden.npz),
 (b) the compiled reference itself (oracle/_ref), when present,
 (c) the reference's published known answers on its own demo data (sba-1.6p/demo/README.txt:131-206)."""
import numpy as np
import pytest
from oracle import ba_port, sba_ref
from argus_gui_b200 import scenes
from helpers import README_KNOWN, have_demo, load_demo, rel, rel_blocks


def test_projection_golden(golden_ba):
    g = golden_ba
    hx = ba_port.port_project(g["G1_p0r"], g["G1_vmask"], g["G1_rot0"])
    assert rel(hx, g["G1_hx"]) < 1e-13


@pytest.mark.parametrize("nk,nd", [(0, 0), (2, 3), (5, 4)])
def test_jacobian_golden(golden_ba, nk, nd):
    g = golden_ba
    A, B = ba_port.port_jacobian(g["G1_p0r"], g["G1_vmask"], g["G1_rot0"], nk, nd)
    rA, rB = g["G1_A_%d%d" % (nk, nd)], g["G1_B_%d%d" % (nk, nd)]
    assert rel(A, rA) < 1e-12 and rel(B, rB) < 1e-12
    assert np.array_equal(A == 0, rA == 0) or nk + nd == 0     # the same columns are zeroed


@pytest.mark.parametrize("it", [1, 3, 8])
def test_trajectory_golden(golden_ba, it):
    g = golden_ba
    p, info, ret = ba_port.port_motstr(g["G1_p0"], g["G1_x"], g["G1_vmask"], g["G1_rot0"], 2, 3, itmax=it)
    rinfo = g["G1_info_it%d" % it]
    assert ret == int(rinfo[5])
    assert list(info[5:]) == list(rinfo[5:])
    assert abs(info[1] - rinfo[1]) <= 1e-10 * rinfo[1]
    assert rel_blocks(p, g["G1_p_it%d" % it], 4) < 1e-8


@pytest.mark.parametrize("it", [2, 6])
def test_rejected_steps_golden(golden_ba, it):
    """sba 1.6's behaviour after a rejected step (diagonals not restored) is reproduced:
    same number of linear systems, same cost."""
    g = golden_ba
    p, info, ret = ba_port.port_motstr(g["G2_p0"], g["G2_x"], g["G2_vmask"], g["G2_rot0"], 0, 0, itmax=it, opts=g["G2_opts"], mcon=1)
    rinfo = g["G2_info_it%d" % it]
    assert list(info[5:]) == list(rinfo[5:])
    assert info[9] > info[5]            # there really were rejected attempts
    assert abs(info[1] - rinfo[1]) <= 1e-8 * rinfo[1]


def test_strict_mode_differs_after_rejection(golden_ba):
    g = golden_ba
    _, info_c, _ = ba_port.port_motstr(g["G2_p0"], g["G2_x"], g["G2_vmask"], g["G2_rot0"], 0, 0, itmax=6, opts=g["G2_opts"], mcon=1)
    _, info_s, _ = ba_port.port_motstr(g["G2_p0"], g["G2_x"], g["G2_vmask"], g["G2_rot0"], 0, 0, itmax=6, opts=g["G2_opts"], mcon=1, compat=False)
    assert info_c[9] != info_s[9] or info_c[1] != info_s[1]


def test_wand_mode_converged_golden(golden_ba):
    g = golden_ba
    p, info, ret = ba_port.port_motstr(g["G3_p0"], g["G3_x"], g["G3_vmask"], g["G3_rot0"], 5, 4, itmax=150)
    assert abs(info[1] - g["G3_info"][1]) <= 1e-8 * g["G3_info"][1]
    # fixed intrinsics really stay fixed
    assert np.array_equal(p[:48].reshape(3, 16)[:, 0:5], g["G3_p0"][:48].reshape(3, 16)[:, 0:5])
    assert np.array_equal(p[:48].reshape(3, 16)[:, 6:10], g["G3_p0"][:48].reshape(3, 16)[:, 6:10])


def test_too_few_measurements_is_an_error():
    s = scenes.make_ba_scene(3, 2, 3, seed=1)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    _, _, ret = ba_port.port_motstr(p0, s["x"], s["vmask"], rot0, 0, 0, itmax=5)
    assert ret == -1


@pytest.mark.skipif(not sba_ref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("cfg", [(4, 300, 3, 2, 3, 0, 1e-3, 1.0), (6, 500, 4, 0, 0, 1, 1e-8, 40.0)])
def test_port_vs_compiled_reference_live(cfg):
    m, n, views, nk, nd, mcon, tau, ps = cfg
    s = scenes.make_ba_scene(m, n, views, seed=99, pert_scale=ps, outlier_frac=0.05 if ps > 1 else 0.0)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    rng = np.random.default_rng(3)
    p0r = p0.copy(); p0r[:16 * m].reshape(m, 16)[:, 10:13] = rng.normal(0, 0.02, size=(m, 3))
    assert rel(ba_port.port_project(p0r, s["vmask"], rot0), sba_ref.ref_project(p0r, s["vmask"], rot0)) < 1e-13
    A, B = ba_port.port_jacobian(p0r, s["vmask"], rot0, nk, nd)
    rA, rB = sba_ref.ref_jacobian(p0r, s["vmask"], rot0, nk, nd)
    assert rel(A, rA) < 1e-12 and rel(B, rB) < 1e-12
    opts = [tau, 1e-12, 1e-12, 1e-12, 0.0]
    for it in (1, 5, 10):
        pr, ir, rr = sba_ref.ref_motstr(p0, s["x"], s["vmask"], rot0, nk, nd, itmax=it, opts=opts, mcon=mcon)
        pp, ip, rp = ba_port.port_motstr(p0, s["x"], s["vmask"], rot0, nk, nd, itmax=it, opts=opts, mcon=mcon)
        assert rr == rp and list(ir[5:]) == list(ip[5:])
        assert abs(ir[1] - ip[1]) <= 1e-8 * ir[1]


@pytest.mark.skipif(not have_demo(), reason="reference demo data (oracle/_ref/demo) not present")
@pytest.mark.parametrize("name", ["7cams", "9cams", "54camsvarKD"])
def test_readme_known_answers(name):
    """Final mean squared reprojection error of the reference's published sample runs, through the KDRTS model."""
    cf, pf, nk, nd, e0, e1, digits = README_KNOWN[name]
    cams, pts, vmask, x = load_demo(cf, pf)
    p0, rot0 = scenes.pack_ba_problem(cams, pts)
    p, info, ret = ba_port.port_motstr(p0, x, vmask, rot0, nk, nd, itmax=150)
    nproj = x.shape[0]
    assert abs(info[0] / nproj - e0) < 5e-5 * e0
    assert abs(info[1] / nproj - e1) < 10.0 ** (-digits) * 2
