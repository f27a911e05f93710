"""sbaArgusDriver mirror: CPU checks of the host logic against golden output of the reference's own driver constructor;
GPU check of fix() (BASELINE config 1: 3-camera wand calibration) against the CPU oracle on the same sba inputs."""
import os
import numpy as np
import pytest

from argus_gui_b200 import scenes, sbaDriver, sba

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("backend", ["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def test_constructor_matches_reference_golden(tmp_path, backend, monkeypatch):
    """Host logic of the constructor (parse / count / rearrange / stacking) + the initial triangulation.  On CPU the
    triangulator is the oracle's restatement (the product one needs the GPU); with a GPU the product path runs as shipped."""
    if backend == "oracle":
        from oracle import tri_port
        monkeypatch.setattr(sbaDriver, "multiTriangulator", tri_port.multiTriangulator)
    g = np.load(os.path.join(ROOT, "tests", "golden", "driver_golden.npz"))
    ppts, uppts, cams = g["ppts"].copy(), g["uppts"].copy(), g["cams"].copy()
    d = sbaDriver.sbaArgusDriver(ppts, uppts, cams, display=False, scale=0.5, modeString="01", name=str(tmp_path / "x"), temp="", report=False)
    assert list(d.order) == list(g["order"]) and d.nppts == int(g["nppts"]) and d.nuppts == int(g["nuppts"])
    assert np.array_equal(d.count(g["ppts"]) + d.count(g["uppts"]), g["counts"])
    # inputs are mutated in place exactly as the reference does (camera swap)
    assert np.array_equal(ppts, g["ppts_after"], equal_nan=True) and np.array_equal(uppts, g["uppts_after"], equal_nan=True)
    assert np.array_equal(cams, g["cams_after"])
    assert d.indices["paired"][0] == list(g["paired0"]) and d.indices["paired"][1] == list(g["paired1"])
    assert d.indices["unpaired"][0] == list(g["unpaired0"])
    assert d.pts.shape == g["pts"].shape
    assert np.array_equal(d.pts[:, 3:], g["pts"][:, 3:], equal_nan=True)          # stacked pixel rows: identical
    assert np.abs(d.pts[:, :3] - g["pts"][:, :3]).max() <= 1e-9 * np.abs(g["pts"][:, :3]).max()   # initial triangulation
    assert np.abs(d.ext - g["ext"]).max() < 1e-9


@pytest.mark.gpu
def test_c1_wand_calibration_end_to_end(tmp_path):
    from oracle import ba_port
    w = scenes.make_wand_scene(nframes=2000, nbackground=100, m=3, seed=1)
    d = sbaDriver.sbaArgusDriver(w["ppts"].copy(), w["uppts"].copy(), w["cams"].copy(), display=False, scale=0.5, modeString="01",
                                 name=str(tmp_path / "c1"), temp=str(tmp_path), report=False)
    assert d.pts.shape[0] > 3500
    res = d.fix()
    # the BA inside fix() equals the oracle run on the same sba inputs
    points = sba.Points.fromDylan(d.pts)
    cameras = sba.Cameras.fromDylan(np.hstack((d.cams, d.ext)))
    a, rot0 = cameras.pack()
    p0 = np.concatenate([a.ravel(), points.B.ravel()])
    po, io, ro = ba_port.port_motstr(p0, points.X, points.vmask, rot0.ravel(), 5, 4, itmax=150)
    assert abs(d.info.finalError - io[1]) <= 1e-6 * io[1]
    cam_o = sba.Cameras.unpack(po[:48].reshape(3, 16), rot0)
    # gauge-invariant: the wand length statistics and the reprojection cost agree; intrinsics stayed fixed, r2 was refined
    assert np.array_equal(d.newcameras.arr[:, 0:5], np.hstack((d.cams, d.ext))[:, 0:5])
    assert np.array_equal(d.newcameras.arr[:, 6:10], d.cams[:, 6:10])
    assert np.abs(d.newcameras.arr[:, 5] - cam_o[:, 5]).max() < 1e-4
    assert isinstance(res.wandscore, float) and res.wandscore < 2.0            # noise 0.5 px -> sub-percent wand length spread
    paired = np.loadtxt(str(tmp_path / "c1-paired-points-xyz.csv"), delimiter=",", skiprows=1)
    ok = ~np.isnan(paired).any(axis=1)
    lengths = np.linalg.norm(paired[ok, :3] - paired[ok, 3:], axis=1)
    assert abs(lengths.mean() - 0.5) < 1e-9                                    # scaled by the wand length (sbaDriver.py:703-713)
    for f in ("c1-sba-profile.txt", "c1-clicker-profile.txt", "c1-dlt-coefficients.csv", "c1-unpaired-points-xyz.csv", "c1-camera-1-profile.txt"):
        assert os.path.exists(str(tmp_path / f)), f
    prof = np.loadtxt(str(tmp_path / "c1-sba-profile.txt"))
    assert prof.shape == (3, 17)
    dl = np.loadtxt(str(tmp_path / "c1-dlt-coefficients.csv"), delimiter=",")
    assert dl.shape == (11, 3) and res.dlterrors.max() < 2.0


@pytest.mark.gpu
def test_c1_with_the_wand_length_inside_the_solver(tmp_path):
    """Extension (parity unpinned): wand_weight > 0 constrains the wand length during the adjustment; the wand score (spread of
    the reconstructed lengths, sbaDriver.py:779-781) drops, the files keep their formats and the scale is still the wand's."""
    scores = {}
    for ww in (0.0, 3000.0):
        w = scenes.make_wand_scene(nframes=600, nbackground=60, m=3, seed=4, noise=1.0)
        d = sbaDriver.sbaArgusDriver(w["ppts"].copy(), w["uppts"].copy(), w["cams"].copy(), display=False, scale=0.5, modeString="01",
                                     name=str(tmp_path / ("w%d" % int(ww))), temp=str(tmp_path), report=False, wand_weight=ww)
        res = d.fix()
        scores[ww] = res.wandscore
        paired = np.loadtxt(str(tmp_path / ("w%d-paired-points-xyz.csv" % int(ww))), delimiter=",", skiprows=1)
        ok = ~np.isnan(paired).any(axis=1)
        assert abs(np.linalg.norm(paired[ok, :3] - paired[ok, 3:], axis=1).mean() - 0.5) < 1e-9
        assert res.dlterrors.max() < 4.0
    assert scores[3000.0] < 0.5 * scores[0.0]
