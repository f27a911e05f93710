"""GPU parity tests of the bundle-adjustment path.  Everything goes through the C ABI (ctypes) and is checked
against the CPU oracle (oracle/ba_port.c, oracle/ba_numpy.py), the golden vectors produced by the reference's own
compiled code, and - at BASELINE sizes - size-independent properties.  Protocol: SURVEY.md appendix D (P0..P6, M0)."""
import threading
import numpy as np
import pytest

from argus_gui_b200 import scenes, ba, _lib
from oracle import ba_port, ba_numpy
from helpers import README_KNOWN, have_demo, load_demo, rel, rel_blocks

pytestmark = pytest.mark.gpu


def test_p0_projection_golden(golden_ba):
    g = golden_ba
    B = ba.BundleAdjuster().upload(g["G1_vmask"], g["G1_p0r"], g["G1_rot0"], g["G1_x"], 0, 0)
    assert rel(B.project(), g["G1_hx"]) < 1e-13          # tolerance: 1e-13 relative (P0)
    B.close()


@pytest.mark.parametrize("cfg", [(4, 500, 3, 2, 3, 0, 0), (8, 2000, 6, 0, 0, 0, 0), (3, 300, 3, 5, 4, 0, 0), (6, 800, 4, 2, 0, 1, 0),
                                 (5, 600, 4, 0, 3, 2, 25),
                                 # the reduced-system solvers by size: 32 unknowns (one diagonal block), 240 (a partial last block), 512 (the
                                 # largest the thread-block-cluster kernel takes), 528 (the cooperative-grid kernel)
                                 (2, 200, 2, 0, 0, 0, 0), (16, 1500, 9, 0, 0, 1, 0), (32, 1500, 10, 0, 0, 0, 0), (33, 1500, 10, 0, 0, 0, 0)])
def test_p2_p3_normal_equations_and_step(cfg):
    """U, ea, V, eb, S, E and one step dp against the numpy restatement fed with the oracle's Jacobian."""
    m, n, views, nk, nd, mcon, ncon = cfg
    s = scenes.make_ba_scene(m, n, views, seed=3 * m + n)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    rng = np.random.default_rng(1)
    p0r = p0.copy(); p0r[:16 * m].reshape(m, 16)[:, 10:13] = rng.normal(0, 0.01, size=(m, 3))
    B = ba.BundleAdjuster().upload(s["vmask"], p0r, rot0, s["x"], nk, nd, ncon=ncon, mcon=mcon)
    A, Bj = ba_port.port_jacobian(p0r, s["vmask"], rot0, nk, nd)
    hx = ba_port.port_project(p0r, s["vmask"], rot0)
    pi, ci = np.nonzero(s["vmask"])
    mu = 0.37
    ne = ba_numpy.normal_equations(A, Bj, s["x"] - hx, pi, ci, n, m, mu, mcon=mcon, ncon=ncon)
    g = B.normal_equations(mu, mcon=mcon)
    V6 = np.stack([ne["V"][:, 0, 0], ne["V"][:, 0, 1], ne["V"][:, 0, 2], ne["V"][:, 1, 1], ne["V"][:, 1, 2], ne["V"][:, 2, 2]], axis=1)
    for name, a, b in (("U", g["U"], ne["U"]), ("ea", g["ea"], ne["ea"]), ("V", g["V"], V6), ("eb", g["eb"], ne["eb"]),
                       ("S", g["S"], ne["S"]), ("E", g["E"], ne["E"])):
        assert rel(a, b) < 1e-11, name                    # tolerance 1e-11 (P2, summation-order noise)
    assert g["scal"][7] == 1.0 and ne["solved"]
    assert rel(g["dp"], ne["dp"]) < 1e-8                   # P3 (includes the Cholesky solve)
    B.close()


@pytest.mark.parametrize("it", [1, 3, 8])
def test_p4_trajectory_golden(golden_ba, it):
    g = golden_ba
    p, info, ret = ba.solve_motstr(g["G1_vmask"], g["G1_p0"], g["G1_rot0"], g["G1_x"], 2, 3, itmax=it)
    rinfo = g["G1_info_it%d" % it]
    assert ret == int(rinfo[5])
    assert list(info[5:]) == list(rinfo[5:])               # iterations, stop reason, nfev, njev, nlss
    assert abs(info[1] - rinfo[1]) <= 1e-10 * rinfo[1]     # cost sequence 1e-10 (P4)
    assert rel_blocks(p, g["G1_p_it%d" % it], 4) < 1e-8    # parameters 1e-8 (P4)
    assert abs(info[0] - rinfo[0]) <= 1e-12 * rinfo[0]


@pytest.mark.parametrize("it", [2, 6])
def test_rejected_steps_golden(golden_ba, it):
    """Same damping trajectory as sba 1.6 through rejected steps (compat mode): identical number of linear systems."""
    g = golden_ba
    p, info, ret = ba.solve_motstr(g["G2_vmask"], g["G2_p0"], g["G2_rot0"], g["G2_x"], 0, 0, itmax=it, opts=g["G2_opts"], mcon=1)
    rinfo = g["G2_info_it%d" % it]
    assert list(info[5:]) == list(rinfo[5:])
    assert info[9] > info[5]
    assert abs(info[1] - rinfo[1]) <= 1e-8 * rinfo[1]
    assert rel_blocks(p, g["G2_p_it%d" % it], 4) < 1e-6


def test_strict_mode_restores_diagonals(golden_ba):
    g = golden_ba
    pc, ic, _ = ba.solve_motstr(g["G2_vmask"], g["G2_p0"], g["G2_rot0"], g["G2_x"], 0, 0, itmax=6, opts=g["G2_opts"], mcon=1, compat=True)
    ps, is_, _ = ba.solve_motstr(g["G2_vmask"], g["G2_p0"], g["G2_rot0"], g["G2_x"], 0, 0, itmax=6, opts=g["G2_opts"], mcon=1, compat=False)
    po, io, _ = ba_port.port_motstr(g["G2_p0"], g["G2_x"], g["G2_vmask"], g["G2_rot0"], 0, 0, itmax=6, opts=g["G2_opts"], mcon=1, compat=False)
    assert list(is_[5:]) == list(io[5:])
    assert abs(is_[1] - io[1]) <= 1e-8 * io[1]
    assert ic[9] != is_[9] or ic[1] != is_[1]


def test_p5_wand_mode_converged_golden(golden_ba):
    g = golden_ba
    p, info, ret = ba.solve_motstr(g["G3_vmask"], g["G3_p0"], g["G3_rot0"], g["G3_x"], 5, 4, itmax=150)
    assert abs(info[1] - g["G3_info"][1]) <= 1e-8 * g["G3_info"][1]       # converged cost 1e-8 (P5)
    c, c0 = p[:48].reshape(3, 16), g["G3_p0"][:48].reshape(3, 16)
    assert np.array_equal(c[:, 0:5], c0[:, 0:5]) and np.array_equal(c[:, 6:10], c0[:, 6:10])   # fixed stay bit-identical
    # gauge-invariant quantity: reprojections at the solution agree with the reference's solution
    s_vmask = g["G3_vmask"]
    hx_g = ba_port.port_project(p, s_vmask, g["G3_rot0"]); hx_r = ba_port.port_project(g["G3_p"], s_vmask, g["G3_rot0"])
    assert np.abs(hx_g - hx_r).max() < 1e-3                                  # pixels


@pytest.mark.parametrize("cfg", [(4, 400, 3, 2, 3, 0, 1e-3, 1.0), (6, 700, 4, 0, 0, 1, 1e-8, 40.0), (16, 3000, 10, 0, 0, 0, 1e-3, 1.0)])
def test_p4_trajectory_vs_oracle_live(cfg):
    m, n, views, nk, nd, mcon, tau, ps = cfg
    s = scenes.make_ba_scene(m, n, views, seed=123, pert_scale=ps, outlier_frac=0.05 if ps > 1 else 0.0)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    opts = [tau, 1e-12, 1e-12, 1e-12, 0.0]
    for it in (1, 4, 10):
        po, io, ro = ba_port.port_motstr(p0, s["x"], s["vmask"], rot0, nk, nd, itmax=it, opts=opts, mcon=mcon)
        pg, ig, rg = ba.solve_motstr(s["vmask"], p0, rot0, s["x"], nk, nd, itmax=it, opts=opts, mcon=mcon)
        assert ro == rg and list(io[5:]) == list(ig[5:])
        assert abs(io[1] - ig[1]) <= 1e-8 * io[1]
        assert rel_blocks(pg, po, m) < 1e-6


@pytest.mark.skipif(not have_demo(), reason="reference demo data (oracle/_ref/demo) not present")
@pytest.mark.parametrize("name", ["7cams", "9cams", "54cams", "54camsvarK", "54camsvarKD"])
def test_p6_readme_known_answers(name):
    """The reference's published sample runs (sba-1.6p/demo/README.txt:131-206) through the CUDA solver."""
    cf, pf, nk, nd, e0, e1, digits = README_KNOWN[name]
    cams, pts, vmask, x = load_demo(cf, pf)
    p0, rot0 = scenes.pack_ba_problem(cams, pts)
    p, info, ret = ba.solve_motstr(vmask, p0, rot0, x, nk, nd, itmax=150)
    nproj = x.shape[0]
    assert ret >= 0
    assert abs(info[0] / nproj - e0) < 5e-5 * e0
    assert abs(info[1] / nproj - e1) < 2e-6                  # 6 digits of the published final error


def test_errors_and_edge_cases():
    s = scenes.make_ba_scene(3, 2, 3, seed=1)                 # fewer measurements than unknowns
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    _, _, ret = ba.solve_motstr(s["vmask"], p0, rot0, s["x"], 0, 0, itmax=5)
    assert ret == _lib.ARGUS_E_SBA
    s = scenes.make_ba_scene(4, 200, 3, seed=2)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    p, info, ret = ba.solve_motstr(s["vmask"], p0, rot0, s["x"], 0, 0, itmax=0)     # itmax = 0: nothing to do
    assert ret == 0 and np.array_equal(p, p0)
    with pytest.raises(ValueError):                             # more cameras than the visibility mask supports
        ba.BundleAdjuster().upload(np.ones((10, 65), dtype=np.int8), np.zeros(16 * 65 + 30), np.zeros(4 * 65), np.zeros((650, 2)))
    # a point behind/at the camera plane makes the cost non-finite: stop reason 7, SBA_ERROR (sba_levmar.c:1289-1295)
    pbad = p0.copy(); pbad[16 * 4:16 * 4 + 3] = np.nan
    _, info, ret = ba.solve_motstr(s["vmask"], pbad, rot0, s["x"], 0, 0, itmax=5)
    assert ret == _lib.ARGUS_E_SBA and info[6] == 7
    # ragged visibility: points seen by 2..m cameras, one camera seeing few points
    rng = np.random.default_rng(5)
    s = scenes.make_ba_scene(6, 500, 6, seed=4)
    vm = s["vmask"].copy()
    keep = np.ones_like(vm)
    for i in range(500):
        nd_ = rng.integers(0, 5)
        keep[i, rng.choice(6, size=nd_, replace=False)] = 0
    keep[20:, 5] = 0
    full_idx = np.cumsum(vm.ravel()).reshape(vm.shape) - 1
    vm2 = vm & keep
    x2 = s["x"][full_idx[vm2 != 0]]
    po, io, ro = ba_port.port_motstr(p0 := scenes.pack_ba_problem(s["cams0"], s["pts0"])[0], x2, vm2, scenes.pack_ba_problem(s["cams0"], s["pts0"])[1], 2, 3, itmax=6)
    pg, ig, rg = ba.solve_motstr(vm2, p0, scenes.pack_ba_problem(s["cams0"], s["pts0"])[1], x2, 2, 3, itmax=6)
    assert ro == rg and list(io[5:]) == list(ig[5:]) and abs(io[1] - ig[1]) <= 1e-9 * io[1]


def test_full_size_properties_c3():
    """BASELINE config 3 (8 cameras, 200k points, 1.2M observations, all 10 intrinsics free): size-independent
    properties - monotone cost, invariance to a permutation of the points, agreement of the resident and one-shot
    interfaces, and recovery of a noise-free scene (round trip)."""
    m, n, views = 8, 200_000, 6
    s = scenes.make_ba_scene(m, n, views, seed=3)
    assert s["nobs"] == 1_200_000
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    opts = [1e-3, 0.0, 0.0, 0.0, 0.0]
    B = ba.BundleAdjuster().upload(s["vmask"], p0, rot0, s["x"], 0, 0)
    costs = []
    for it in (1, 2, 4):
        B.reset()
        ret, info = B.solve(itmax=it, opts=opts)
        assert ret == it and info[5] == it
        costs.append(info[1])
    assert info[0] > costs[0] > costs[1] > costs[2]
    p_res = B.download()
    B.close()
    p_one, info_one, _ = ba.solve_motstr(s["vmask"], p0, rot0, s["x"], 0, 0, itmax=4, opts=opts)
    assert np.array_equal(p_res, p_one)                      # deterministic (no floating-point atomics): bit-identical run to run
    # permutation of the points
    perm = np.random.default_rng(0).permutation(n)
    vm_p = s["vmask"][perm]
    obs_pt = np.repeat(np.arange(n), s["vmask"].sum(axis=1).astype(np.int64))
    newpos = np.empty(n, dtype=np.int64); newpos[perm] = np.arange(n)
    key = newpos[obs_pt]
    o = np.argsort(key, kind="stable")
    x_p = s["x"][o]
    p0_p = np.concatenate([p0[:16 * m], p0[16 * m:].reshape(n, 3)[perm].ravel()])
    p_perm, info_perm, _ = ba.solve_motstr(vm_p, p0_p, rot0, x_p, 0, 0, itmax=4, opts=opts)
    assert abs(info_perm[1] - info_one[1]) <= 1e-9 * info_one[1]
    assert rel_blocks(np.concatenate([p_perm[:16 * m], p_perm[16 * m:].reshape(n, 3)[newpos].ravel()]), p_one, m) < 1e-7
    # noise-free round trip: perturb the truth, solve, the reprojection error goes to ~0
    s0 = scenes.make_ba_scene(m, 50_000, views, seed=4, noise=0.0)
    p0, rot0 = scenes.pack_ba_problem(s0["cams0"], s0["pts0"])
    p, info, ret = ba.solve_motstr(s0["vmask"], p0, rot0, s0["x"], 0, 0, itmax=50)
    assert info[1] / s0["nobs"] < 1e-12


def test_full_size_properties_c4():
    """BASELINE config 4 (16 cameras, 2M points, 20M observations - the bench workload): the projection of every observation
    against a numpy recomputation of the cost, a random sample of projections and of the trajectory against the oracle,
    a cost that more than halves in three iterations, bit-identical repetition."""
    m, n, views = 16, 2_000_000, 10
    s = scenes.make_ba_scene(m, n, views, seed=4, cam_seed=4)
    assert s["nobs"] == 20_000_000
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    opts = [1e-3, 0.0, 0.0, 0.0, 0.0]
    B = ba.BundleAdjuster().upload(s["vmask"], p0, rot0, s["x"], 0, 0)
    hx = B.project()
    cost0 = float(((s["x"] - hx) ** 2).sum())
    ret, info = B.solve(itmax=3, opts=opts)
    assert ret == 3 and abs(info[0] - cost0) <= 1e-12 * cost0 and info[1] < 0.5 * info[0]
    p1 = B.download()
    B.reset(); ret2, info2 = B.solve(itmax=3, opts=opts)
    assert np.array_equal(B.download(), p1) and list(info2) == list(info)
    B.close()
    # the first 3000 points with all cameras: their projections are the oracle's
    ns = 3000; no = int(s["vmask"][:ns].sum())
    hx_o = ba_port.port_project(np.concatenate([p0[:16 * m], p0[16 * m:16 * m + 3 * ns]]), s["vmask"][:ns], rot0)
    assert rel(hx[:no], hx_o) < 1e-13
    # a sub-problem small enough for the oracle follows the oracle's trajectory (same kernels, same tile shapes)
    xs = s["x"][:no]
    ps = np.concatenate([p0[:16 * m], p0[16 * m:16 * m + 3 * ns]])
    po, io, ro = ba_port.port_motstr(ps, xs, s["vmask"][:ns], rot0, 0, 0, itmax=3, opts=opts)
    pg, ig, rg = ba.solve_motstr(s["vmask"][:ns], ps, rot0, xs, 0, 0, itmax=3, opts=opts)
    assert ro == rg and list(io[5:]) == list(ig[5:]) and abs(io[1] - ig[1]) <= 1e-9 * io[1]


@pytest.mark.parametrize("mode", ["motstr", "mot", "str"])
def test_m0_two_shards_equal_single(mode):
    """Multi-rank path on one GPU: two solver contexts (threads) exchange through the all-reduce hook; the result
    equals the single-context run to rounding (M0) - for the full, the motion-only and the structure-only driver."""
    import torch
    s = scenes.make_ba_scene(6, 2000, 4, seed=77)
    m, n = 6, 2000
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    B0 = ba.BundleAdjuster(mode=mode).upload(s["vmask"], p0, rot0, s["x"], 2, 3)
    ret_ref, info_ref = B0.solve(itmax=8 if mode != "str" else 3)
    p_ref = B0.download(); B0.close()
    world = 2
    rowptr = np.concatenate([[0], np.cumsum(s["vmask"].sum(axis=1).astype(np.int64))])
    barrier = threading.Barrier(world)
    slots = [None] * world
    results = [None] * world

    def make_hook(rank):
        def hook(ptr, count, op, stream):
            iface = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False), "version": 3, "strides": None}

            class W:
                __cuda_array_interface__ = iface
            t = torch.as_tensor(W(), device="cuda:0")
            torch.cuda.synchronize()
            slots[rank] = t
            barrier.wait()
            tot = slots[0].clone()
            for r in range(1, world):
                tot = tot + slots[r] if op == 0 else torch.maximum(tot, slots[r])
            torch.cuda.synchronize()
            barrier.wait()
            t.copy_(tot)
            torch.cuda.synchronize()
            barrier.wait()
            return 0
        return hook

    def run(rank):
        lo, hi = ba.shard_points(n, world, rank)
        B = ba.BundleAdjuster(mode=mode)
        B.set_comm(rank, world, make_hook(rank))
        p_loc = np.concatenate([p0[:16 * m], p0[16 * m + 3 * lo:16 * m + 3 * hi]])
        B.upload(s["vmask"][lo:hi], p_loc, rot0, s["x"][rowptr[lo]:rowptr[hi]], 2, 3)
        ret, info = B.solve(itmax=8 if mode != "str" else 3)
        results[rank] = (ret, info, B.download(), lo, hi)
        B.close()

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in th]; [t.join() for t in th]
    p_sh = np.zeros_like(p_ref)
    for ret, info, p_loc, lo, hi in results:
        assert ret == ret_ref and list(info[5:]) == list(info_ref[5:])
        assert abs(info[1] - info_ref[1]) <= 1e-11 * info_ref[1]
        p_sh[:16 * m] = p_loc[:16 * m]
        p_sh[16 * m + 3 * lo:16 * m + 3 * hi] = p_loc[16 * m:]
    assert rel_blocks(p_sh, p_ref, m) < 1e-9
    if mode == "mot":
        assert np.array_equal(p_sh[16 * m:], p0[16 * m:])
    if mode == "str":
        assert np.array_equal(p_sh[:16 * m], p0[:16 * m])


def test_config5_shape_per_camera_fixed_intrinsics_and_outliers():
    """BASELINE config 5 at reduced size: 32 cameras, 12 views per point, 10 % outlier tracks, cameras 0-15 with fixed
    intrinsics + distortion, cameras 16-31 fully free (per-camera extension, argus_ba_set_camera_fix) - against the oracle
    port with the same per-camera counts."""
    m, n, views = 32, 6000, 12
    s = scenes.make_ba_scene(m, n, views, seed=5, outlier_frac=0.1, radius=9.0)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    nk = np.array([5] * 16 + [0] * 16); nd = np.array([5] * 16 + [0] * 16)
    B = ba.BundleAdjuster().upload(s["vmask"], p0, rot0, s["x"], 0, 0)
    B.set_camera_fix(nk, nd)
    ret, info = B.solve(itmax=6)
    pg = B.download()
    B.close()
    po, io, ro = ba_port.port_motstr(p0, s["x"], s["vmask"], rot0, 0, 0, itmax=6, cam_fix=(nk, nd))
    assert ret == ro and list(info[5:]) == list(io[5:])
    assert abs(info[1] - io[1]) <= 1e-8 * io[1]
    assert rel_blocks(pg, po, m) < 1e-6
    cg = pg[:16 * m].reshape(m, 16); c0 = p0[:16 * m].reshape(m, 16)
    assert np.array_equal(cg[:16, :10], c0[:16, :10])         # fixed cameras: intrinsics and distortion untouched
    assert np.abs(cg[16:, 0] - c0[16:, 0]).min() > 0            # free cameras: refined


def _subset(s, keep):
    """Scene restricted to the visibility entries where ``keep`` is non-zero: (vmask, x)."""
    vm = s["vmask"]
    idx = np.cumsum(vm.ravel()).reshape(vm.shape) - 1
    vm2 = (vm != 0) & (keep != 0)
    return vm2.astype(np.uint8), s["x"][idx[vm2]]


def test_unobserved_points_and_cameras():
    """Rows / columns of the visibility mask that are entirely zero (a point nobody sees, a camera that sees nothing): their
    blocks are just mu I (sba_levmar.c:992-1001), parameters must not move, the rest must follow the oracle."""
    m, n = 5, 600
    s = scenes.make_ba_scene(m, n, 4, seed=12)
    keep = np.ones_like(s["vmask"])
    keep[[0, 17, 18, 599]] = 0            # unobserved points, first / last / adjacent
    keep[:, 2] = 0                        # camera 2 sees nothing
    vm, x = _subset(s, keep)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    po, io, ro = ba_port.port_motstr(p0, x, vm, rot0, 2, 3, itmax=6)
    pg, ig, rg = ba.solve_motstr(vm, p0, rot0, x, 2, 3, itmax=6)
    assert ro == rg and list(io[5:]) == list(ig[5:]) and abs(io[1] - ig[1]) <= 1e-9 * io[1]
    assert rel_blocks(pg, po, m) < 1e-7
    assert np.array_equal(pg[32:48], p0[32:48])                               # camera 2: zero gradient, no motion
    for i in (0, 17, 18, 599):
        assert np.array_equal(pg[16 * m + 3 * i:16 * m + 3 * i + 3], p0[16 * m + 3 * i:16 * m + 3 * i + 3])


def test_sixty_four_cameras():
    """The visibility bit mask holds at most 64 cameras: the largest supported camera count against the oracle."""
    m, n = 64, 700
    s = scenes.make_ba_scene(m, n, 9, seed=64)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    po, io, ro = ba_port.port_motstr(p0, s["x"], s["vmask"], rot0, 5, 5, itmax=3)
    pg, ig, rg = ba.solve_motstr(s["vmask"], p0, rot0, s["x"], 5, 5, itmax=3)
    assert ro == rg and list(io[5:]) == list(ig[5:]) and abs(io[1] - ig[1]) <= 1e-9 * io[1]
    assert rel_blocks(pg, po, m) < 1e-7


@pytest.mark.parametrize("ncon,mcon", [(400, 0), (0, 3), (399, 3)])
def test_almost_everything_fixed(ncon, mcon):
    """ncon = n (every point fixed), mcon = m - 1 (one free camera) and both at once."""
    m, n = 4, 400
    s = scenes.make_ba_scene(m, n, 4, seed=33)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    po, io, ro = ba_port.port_motstr(p0, s["x"], s["vmask"], rot0, 0, 0, itmax=5, ncon=ncon, mcon=mcon)
    pg, ig, rg = ba.solve_motstr(s["vmask"], p0, rot0, s["x"], 0, 0, itmax=5, ncon=ncon, mcon=mcon)
    assert ro == rg and list(io[5:]) == list(ig[5:]) and abs(io[1] - ig[1]) <= 1e-9 * io[1]
    assert rel_blocks(pg, po, m) < 1e-7
    assert np.array_equal(pg[:16 * mcon], p0[:16 * mcon]) and np.array_equal(pg[16 * m:16 * m + 3 * ncon], p0[16 * m:16 * m + 3 * ncon])


@pytest.mark.parametrize("stop_at", [0, 2, 4])
def test_gradient_stop_matches_oracle(golden_ba, stop_at):
    """Stop reason 1 (||J^T e||_inf <= eps1, sba_levmar.c:977-981) at the first linearisation and at later ones - where the
    solver has already queued the first damping attempt behind the linearisation and must discard it: iteration count,
    function / Jacobian evaluations and linear systems must be the reference's (no extra attempt counted)."""
    g = golden_ba
    args = (g["G1_p0"], g["G1_x"], g["G1_vmask"], g["G1_rot0"], 2, 3)
    # gradient norm the oracle sees at the linearisation of iteration `stop_at` = info[2] of a run stopped by itmax there
    _, iprobe, _ = ba_port.port_motstr(*args, itmax=stop_at + 1)
    if stop_at == 0:
        eps1 = 1e30
    else:
        _, ia, _ = ba_port.port_motstr(*args, itmax=stop_at)          # info[2]: gradient at the last linearisation done
        _, ib, _ = ba_port.port_motstr(*args, itmax=stop_at + 1)
        assert ib[2] < ia[2]
        eps1 = 0.5 * (ia[2] + ib[2])                                   # iteration stop_at's linearisation is the first below it
    opts = [1e-3, eps1, 1e-12, 1e-12, 0.0]
    po, io, ro = ba_port.port_motstr(*args, itmax=50, opts=opts)
    assert int(io[6]) == 1
    pg, ig, rg = ba.solve_motstr(g["G1_vmask"], g["G1_p0"], g["G1_rot0"], g["G1_x"], 2, 3, itmax=50, opts=opts)
    assert rg == ro and list(ig[5:]) == list(io[5:]) and abs(ig[1] - io[1]) <= 1e-10 * io[1] and abs(ig[2] - io[2]) <= 1e-6 * io[2]
    assert rel_blocks(pg, po, 4) < 1e-8
