"""fix()-level parity of the wand-calibration driver against golden outputs of the REFERENCE's own driver
(tests/golden/fix_golden.npz, produced by tests/golden/make_fix_golden.py: /root/reference/argus_gui/sbaDriver.py run end to
end with the bundle adjustment done by the reference's compiled C code).  One case per alignment mode: no reference points,
1- / 2- / 3- / 4-point axis references, gravity, plane; mode strings '01', '00', '12', '10' (nccalib 5 / 4, ncdist 5 / 4 / 3).

* CPU: the driver's host logic (stacking, camera re-ordering, wand scale and score, transform(), DLT sign change, outlier
  report, every output file) with the numerical back ends replaced by the CPU oracles (initial triangulation, bundle
  adjustment - the reference's C code itself when oracle/_ref is built -, undistortion, DLT fit).
* GPU: the product path as shipped (CUDA triangulation, bundle adjustment, undistortion, DLT fit), compared gauge-invariantly
  at the north-star tolerance of 1e-6 wherever the damping trajectory is the reference's (see the test for the other case)."""
import os
import numpy as np
import pytest

from argus_gui_b200 import sbaDriver, ba, sba
from argus_gui_b200 import dlt as dlt_host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "fix_golden.npz"))
CASES = [str(c) for c in G["cases"]]


def _cpu_backends(monkeypatch):
    from oracle import tri_port, dlt_port, sba_ref, ba_port
    monkeypatch.setattr(sbaDriver, "multiTriangulator", tri_port.multiTriangulator)

    def solve_motstr(vmask, p, rot0, x, nccalib=0, ncdist=0, itmax=150, opts=None, ncon=0, mcon=0, verbose=0, compat=True, inplace=False):
        fn = sba_ref.ref_motstr if sba_ref.available() else ba_port.port_motstr
        return fn(p, x, vmask, rot0, nccalib, ncdist, itmax=itmax, opts=list(opts) if opts is not None else None, ncon=ncon, mcon=mcon)
    monkeypatch.setattr(ba, "solve_motstr", solve_motstr)
    monkeypatch.setattr(sbaDriver, "undistort_pts", lambda pts, prof: dlt_port.undistort_pts_f32(pts, np.asarray(prof)[[0, 1, 2, 5, 6, 7, 8, 9]]))

    def fit(xyz, uv, both_coordinates=False):
        L, e, idx = dlt_port.solve_dlt(xyz, uv, both_coordinates=both_coordinates)
        err = np.full(xyz.shape[0], np.nan); err[idx] = e
        return L, err, (len(idx), float(e.sum()), float(e.mean()), float(e.std()))
    monkeypatch.setattr(dlt_host, "fit", fit)


def _run(case, tmp_path, same_start=False):
    g = lambda k: G[case + "_" + k]
    ref = g("ref").copy() if (case + "_ref") in G.files else None
    d = sbaDriver.sbaArgusDriver(g("ppts").copy(), g("uppts").copy(), g("cams").copy(), display=False, scale=0.5, modeString=str(g("mode")),
                                 ref=ref, name=str(tmp_path / case), temp=str(tmp_path), report=False, outputCPs=True,
                                 reference_type=str(g("rtype")), recording_frequency=100)
    assert np.abs(d.pts[:, :3] - g("pts")[:, :3]).max() <= 1e-8 * np.abs(g("pts")[:, :3]).max()      # initial structure
    assert np.abs(d.ext - g("ext")).max() <= 1e-8
    if same_start:          # hand the bundle adjustment the reference driver's start to the last bit
        d.pts = g("pts").copy(); d.ext = g("ext").copy()
    res = d.fix()
    return d, res, g


def _pairwise(X, k=60):
    idx = np.linspace(0, X.shape[0] - 1, k).astype(int)
    P = X[idx]
    return np.linalg.norm(P[:, None, :] - P[None, :, :], axis=2)


def _check_outputs(case, d, res, g, tmp_path, tol, aligned):
    rel = lambda a, b: float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))
    assert list(d.order) == list(g("order"))
    assert d.pts.shape == g("pts").shape and np.array_equal(d.pts[:, 3:], g("pts")[:, 3:], equal_nan=True)
    assert abs(res.wandscore - float(g("wandscore"))) <= tol * float(g("wandscore"))          # wand score: gauge invariant
    assert rel(res.dlterrors, g("dlterrors")) <= tol                                           # per-camera DLT rmse: gauge invariant
    assert rel(_pairwise(res.xyzs), _pairwise(g("xyzs"))) <= tol                               # scaled structure up to a rigid motion
    # the 3-sigma report names the same observations
    assert sorted(set(res.index)) == list(g("outlier_index"))
    assert [o["frame"] for o in res.outliers] == list(g("outlier_frames")) and [o["camera"] for o in res.outliers] == list(g("outlier_cams"))
    assert rel([o["error"] for o in res.outliers], g("outlier_err")) <= 10 * tol
    assert np.abs(np.array([o["uv"] for o in res.outliers]) - g("outlier_uv")).max() < 2e-3    # float32 undistortion, redistorted in double
    pre = str(tmp_path / case)
    prof = np.loadtxt(pre + "-sba-profile.txt")
    assert prof.shape == g("sba_profile").shape
    assert rel(prof[:, :10], g("sba_profile")[:, :10]) <= tol                                   # intrinsics + distortion: gauge invariant
    assert rel(np.loadtxt(pre + "-clicker-profile.txt"), g("clicker_profile")) <= tol
    assert np.array_equal(np.loadtxt(pre + "-camera-1-profile.txt").shape, g("camera1_profile").shape)
    pc = np.genfromtxt(pre + "-paired-points-xyz.csv", delimiter=",", skip_header=1)
    uc = np.genfromtxt(pre + "-unpaired-points-xyz.csv", delimiter=",", skip_header=1)
    assert np.array_equal(np.isnan(pc), np.isnan(g("paired_csv"))) and np.array_equal(np.isnan(uc), np.isnan(g("unpaired_csv")))
    ok = ~np.isnan(pc).any(axis=1)
    assert abs(np.linalg.norm(pc[ok, :3] - pc[ok, 3:], axis=1).mean() - 0.5) < 1e-9              # scaled by the wand length
    if aligned:
        # reference points pin the frame: coordinates, DLT coefficients and the written files are comparable as they are
        scale = np.abs(g("xyzs")).max()
        sgn = np.ones(3)
        if case == "plane":
            # the horizontal-plane alignment takes its two in-plane axes from numpy's SVD of the reference points
            # (sbaDriver.py:940-975) and never fixes their signs (its `det == -1` test compares floats exactly): a 1e-12 change of
            # the input can flip x or y.  +z is decided by the side the data lies on and must agree.
            sgn = np.sign((res.xyzs * g("xyzs")).sum(axis=0))
            assert sgn[2] == 1.0
        lsgn = np.array([sgn[0], sgn[1], sgn[2], 1.0, sgn[0], sgn[1], sgn[2], 1.0, sgn[0], sgn[1], sgn[2]])     # L1..L11 multiply x, y, z, 1
        assert np.abs(res.xyzs * sgn - g("xyzs")).max() <= tol * scale
        assert rel(res.dlts[:, :, 0] * lsgn, g("dlts")) <= 20 * tol
        assert rel(np.loadtxt(pre + "-dlt-coefficients.csv", delimiter=",") * lsgn[:, None], g("dlt_csv")) <= 20 * tol
        assert np.nanmax(np.abs(pc * np.tile(sgn, pc.shape[1] // 3) - g("paired_csv"))) <= tol * scale
        assert np.nanmax(np.abs(uc * np.tile(sgn, uc.shape[1] // 3) - g("unpaired_csv"))) <= tol * scale


@pytest.mark.parametrize("case", CASES)
def test_fix_host_logic_matches_reference_driver(case, tmp_path, monkeypatch):
    """CPU: same numbers in, same files out as the reference's driver (the BA runs in the reference's own C code)."""
    from oracle import sba_ref
    _cpu_backends(monkeypatch)
    d, res, g = _run(case, tmp_path, same_start=True)
    tight = sba_ref.available()           # with the port instead of the reference's binaries only the trajectory-level tolerance holds
    tol = 1e-9 if tight else 1e-6
    assert np.abs(d.newcameras.arr - g("newcams")).max() <= tol * np.abs(g("newcams")).max()
    assert np.abs(d.newpoints.B - g("newpts")).max() <= tol * np.abs(g("newpts")).max()
    _check_outputs(case, d, res, g, tmp_path, tol, aligned=(case != "noref"))


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_fix_on_gpu_matches_reference_driver(case, tmp_path):
    """GPU: the driver as shipped against the reference driver's outputs, 1e-6 relative on everything that does not depend on
    the gauge the solver is free to drift in (with reference points the alignment removes that freedom entirely)."""
    # the bundle adjustment starts from the reference driver's own initial structure (the CUDA triangulation agrees with it to
    # 1e-8, asserted in _run; handing over the identical start keeps the two damping trajectories comparable step for step)
    d, res, g = _run(case, tmp_path, same_start=True)
    # Argus never pins the gauge (mcon = 0, the reference camera is last): the reduced system is singular up to the damping, and
    # whether its Cholesky factorisation fails in a given attempt - hence the rest of the damping trajectory and the iteration at
    # which the noise-triggered stop 4 fires - is decided by rounding (the reference does not reproduce its own
    # iteration counts across BLAS thread counts).  Where the trajectory is the reference's (same counters) everything
    # gauge-invariant agrees to 1e-6; otherwise both runs sit in the same minimum, to within where the stopping test left them.
    same_trajectory = list(d.info.info[5:]) == list(g("info")[5:])
    tol = 1e-6 if same_trajectory else 3e-5
    assert abs(d.info.finalError - g("info")[1]) <= (1e-9 if same_trajectory else 1e-6) * g("info")[1]
    assert abs(d.info.initialError - g("info")[0]) <= 1e-12 * g("info")[0]
    _check_outputs(case, d, res, g, tmp_path, tol, aligned=(case != "noref"))


def test_align_to_reference_properties():
    """transform(): the aligned reference points sit where the mode says they should."""
    rng = np.random.default_rng(3)
    Q = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    if np.linalg.det(Q) < 0:
        Q[:, 2] *= -1
    o = np.array([1.0, -2.0, 0.5])
    cloud = rng.normal(size=(50, 3))
    # four axis points: origin, +x, +y, +z of a rotated, shifted frame
    refs = np.vstack([o, o + 2 * Q[:, 0], o + 3 * Q[:, 1], o + 1.5 * Q[:, 2]])
    out = sbaDriver.align_to_reference(np.vstack([refs, cloud]), 4, 'Axis points')
    assert np.abs(out[:4] - np.array([[0, 0, 0], [2, 0, 0], [0, 3, 0], [0, 0, 1.5]])).max() < 1e-12
    assert np.abs(_pairwise(out, 20) - _pairwise(np.vstack([refs, cloud]), 20)).max() < 1e-12       # rigid
    # two points: origin, +z
    out = sbaDriver.align_to_reference(np.vstack([refs[[0, 3]], cloud]), 2, 'Axis points')
    assert np.abs(out[1] - [0, 0, 1.5]).max() < 1e-12
    # three points: the reference builds its fourth point as -(x cross y), which makes source and target frames of opposite
    # handedness; its least-squares ROTATION is then a compromise (pinned by the 'axis3' golden case) - here only: rigid, origin kept
    out = sbaDriver.align_to_reference(np.vstack([refs[:3], cloud]), 3, 'Axis points')
    assert np.abs(out[0]).max() < 1e-12
    assert np.abs(_pairwise(out, 20) - _pairwise(np.vstack([refs[:3], cloud]), 20)).max() < 1e-12
    # gravity: a free-fall track accelerating along -Q[:, 2] ends up accelerating along -z
    t = np.arange(30) / 100.0
    track = o + np.outer(t, [0.3, 0.1, 0.2]) - 0.5 * 9.81 * np.outer(t * t, Q[:, 2])
    out = sbaDriver.align_to_reference(np.vstack([track, cloud]), 30, 'Gravity', 100)
    acc = np.array([np.polyfit(t, out[:30, a], 2)[0] * 2 for a in range(3)])
    assert np.abs(acc - [0, 0, -9.81]).max() < 1e-8
    # plane: reference points in a tilted plane end up in z = 0, centred, the cloud above it
    ab = rng.uniform(-1, 1, size=(15, 2))
    plane = o + ab[:, :1] * Q[:, 0] + ab[:, 1:] * Q[:, 1]
    above = o + rng.uniform(-1, 1, size=(40, 1)) * Q[:, 0] + rng.uniform(0.5, 2, size=(40, 1)) * Q[:, 2]
    out = sbaDriver.align_to_reference(np.vstack([plane, above]), 15, 'Plane')
    assert np.abs(out[:15, 2]).max() < 1e-12 and np.abs(out[:15].mean(axis=0)).max() < 1e-12 and out[15:, 2].min() > 0.4


def test_redistort_matches_opencv():
    import cv2
    from argus_gui_b200 import tools
    rng = np.random.default_rng(5)
    prof = np.array([1000.0, 959.5, 539.5, -0.05, 0.01, 0.001, -0.001, 0.002])
    und = np.column_stack([rng.uniform(100, 1800, 200), rng.uniform(100, 1000, 200)])
    mine = tools.redistort_pts(und.copy(), prof)
    n = np.column_stack([(und[:, 0] - prof[1]) / prof[0], (und[:, 1] - prof[2]) / prof[0], np.ones(200)])
    K = np.array([[prof[0], 0, prof[1]], [0, prof[0], prof[2]], [0, 0, 1.0]])
    cvp = cv2.projectPoints(n, np.eye(3), np.zeros(3), K, prof[-5:])[0].reshape(-1, 2)
    assert np.abs(mine - cvp).max() < 1e-9
