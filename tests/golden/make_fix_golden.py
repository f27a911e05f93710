"""Generates tests/golden/fix_golden.npz by running the REFERENCE's own wand-calibration driver end to end -
sbaArgusDriver.__init__ + fix() + OutlierWindow.buildData / transform / getCoefficients / WandOutputter
(/root/reference/argus_gui/sbaDriver.py:50-415, 676-1165, output.py:29-66) - with the bundle adjustment done by the
reference's compiled C code (oracle/_ref) behind the `sba` names (tests/golden/ref_import.py).  One case per way Argus can
align the result: no reference points, 1- / 2- / 3- / 4-point axis references, gravity, horizontal plane.
Run in the build container: python tests/golden/make_fix_golden.py"""
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import load_reference, ROOT
ref = load_reference("reference")
# the reference's gravity branch uses `scipy.spatial.transform` without a module-level import (sbaDriver.py:929 - it is only
# imported inside updateGraph): supply the name so that the branch can run at all
import scipy.spatial.transform
ref["sbaDriver"].scipy = scipy
from argus_gui_b200 import scenes


def reference_points(kind, w, rng):
    """Pixel rows (k x 2m) of the reference points of each alignment mode, seen by every camera."""
    cams = w["cams_true"]; m = cams.shape[0]
    Rw = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    if np.linalg.det(Rw) < 0:
        Rw[:, 2] *= -1
    o = np.array([0.2, -0.1, 0.3])
    if kind.startswith("axis"):
        k = int(kind[4:])
        X = np.vstack([o, o + 0.8 * Rw[:, 0], o + 0.6 * Rw[:, 1], o + 0.7 * Rw[:, 2]])[:k]
    elif kind == "gravity":
        t = np.arange(25) / 100.0
        g = -9.81 * Rw[:, 2] * 0.05            # the scene is in arbitrary units before the wand scale; any constant acceleration will do
        X = o + np.outer(t, [0.3, 0.1, 0.2]) + 0.5 * np.outer(t * t, g)
    else:
        ab = rng.uniform(-1, 1, size=(12, 2))
        X = o + ab[:, :1] * Rw[:, 0] + ab[:, 1:] * Rw[:, 1] - 1.2 * Rw[:, 2]
    out = np.empty((X.shape[0], 2 * m))
    for j in range(m):
        out[:, 2 * j:2 * j + 2] = scenes.project_points(cams, X, np.full(X.shape[0], j)) + rng.normal(0, 0.2, size=(X.shape[0], 2))
    return out


CASES = [("noref", None, "Axis points", "01"), ("axis1", "axis1", "Axis points", "01"), ("axis2", "axis2", "Axis points", "00"),
         ("axis3", "axis3", "Axis points", "12"), ("axis4", "axis4", "Axis points", "01"), ("gravity", "gravity", "Gravity", "01"),
         ("plane", "plane", "Plane", "10")]
out = {"cases": np.array([c[0] for c in CASES])}
tmp = tempfile.mkdtemp()
# remember the info vector of the bundle adjustment the reference driver runs (it is a local of fix())
from argus_gui_b200 import sba as _sba
_last = {}
_orig_sba = _sba.SparseBundleAdjust
def _capture(cameras, points, options):
    r = _orig_sba(cameras, points, options)
    _last["info"] = np.array(r[2].info)
    return r
_sba.SparseBundleAdjust = _capture
for name, kind, rtype, mode in CASES:
    rng = np.random.default_rng(100 + len(name))
    w = scenes.make_wand_scene(nframes=260, nbackground=40, m=3 if name != "plane" else 4, seed=11 + len(name))
    refpx = reference_points(kind, w, rng) if kind else None
    # one gross outlier so that the 3-sigma report is not empty
    w["uppts"][5, 0:2] += 60.0
    pre = os.path.join(tmp, name)
    d = ref["sbaDriver"].sbaArgusDriver(w["ppts"].copy(), w["uppts"].copy(), w["cams"].copy(), display=False, scale=0.5, modeString=mode,
                                        ref=None if refpx is None else refpx.copy(), name=pre, temp=tmp, report=False, outputCPs=True,
                                        reference_type=rtype, recording_frequency=100)
    d.fix()
    g = d.grapher
    o = {"ppts": w["ppts"], "uppts": w["uppts"], "cams": w["cams"], "mode": np.array(mode), "rtype": np.array(rtype), "order": np.asarray(d.order),
         "pts": d.pts, "ext": d.ext, "info": _last["info"], "newcams": np.loadtxt(os.path.join(tmp, g.key + "_cn.txt")), "newpts": np.loadtxt(os.path.join(tmp, g.key + "_np.txt")),
         "xyzs": g.xyzs, "dlts": np.asarray(g.dlts)[:, :, 0], "dlterrors": g.dlterrors, "wandscore": np.array(float(g.wandscore)),
         "outlier_index": np.array(sorted(set(g.index))), "outlier_frames": np.array([q["frame"] for q in g.outliers]),
         "outlier_cams": np.array([q["camera"] for q in g.outliers]), "outlier_err": np.array([q["error"] for q in g.outliers]),
         "outlier_uv": np.array([q["uv"] for q in g.outliers]).reshape(-1, 2),
         "sba_profile": np.loadtxt(pre + "-sba-profile.txt"), "clicker_profile": np.loadtxt(pre + "-clicker-profile.txt"),
         "dlt_csv": np.loadtxt(pre + "-dlt-coefficients.csv", delimiter=","),
         "paired_csv": np.genfromtxt(pre + "-paired-points-xyz.csv", delimiter=",", skip_header=1),
         "unpaired_csv": np.genfromtxt(pre + "-unpaired-points-xyz.csv", delimiter=",", skip_header=1),
         "camera1_profile": np.loadtxt(pre + "-camera-1-profile.txt")}
    if refpx is not None:
        o["ref"] = refpx
    for k, v in o.items():
        out[name + "_" + k] = v
    print(name, "wand score %.6f" % float(g.wandscore), "dlt errors", g.dlterrors, "outliers", len(g.outliers), "order", d.order)
np.savez_compressed(os.path.join(HERE, "fix_golden.npz"), **out)
print("saved", os.path.getsize(os.path.join(HERE, "fix_golden.npz")), "bytes")
