"""Generates tests/golden/dlt_golden.npz by importing the REFERENCE's own argus_gui/tools.py (needs /root/reference,
cv2, scipy; `argus.ocam` - an external package only used by the omnidirectional branch - is stubbed).
Run in the build container:  python tests/golden/make_dlt_golden.py
"""
import os, sys, types, importlib.util
import numpy as np

REF = os.environ.get("ARGUS_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from argus_gui_b200 import scenes

argus = types.ModuleType("argus"); ocam = types.ModuleType("argus.ocam")
ocam.PointUndistorter = type("PointUndistorter", (), {}); ocam.ocam_model = type("ocam_model", (), {})
argus.ocam = ocam
sys.modules["argus"] = argus; sys.modules["argus.ocam"] = ocam
spec = importlib.util.spec_from_file_location("ref_tools", os.path.join(REF, "argus_gui", "tools.py"))
tools = importlib.util.module_from_spec(spec); spec.loader.exec_module(tools)

out = {}
# case A: C2-like, 4 cameras, 3 tracks, 10 % NaN + hand-made edge cases
s = scenes.make_dlt_scene(400, m=4, ntracks=3, seed=2, nan_frac=0.1)
pts = s["pts"].copy()
m = 4
pts[0, :] = np.nan                       # nothing visible
pts[1, 2:8] = np.nan                     # one camera only (track 0)
pts[2, 4:8] = np.nan                     # exactly two cameras
pts[3, 1] = np.nan                       # u present, v NaN (get_repo_errors checks only u)
pts[4, 6:8] = np.nan                     # last camera missing -> v-equation comes from camera 2
prof = [list(r) for r in s["prof"]]
xyz = np.hstack([tools.uv_to_xyz(pts[:, t * 2 * m:(t + 1) * 2 * m], prof, s["dlt"]) for t in range(3)])
err = tools.get_repo_errors(xyz, pts, prof, s["dlt"])
out.update(A_pts=pts, A_prof=s["prof"], A_dlt=s["dlt"], A_xyz=xyz, A_err=err)
# case B: 6 cameras, stronger distortion, 2 tracks, 30 % NaN
s = scenes.make_dlt_scene(300, m=6, ntracks=2, seed=5, nan_frac=0.3)
prof = s["prof"].copy(); prof[:, 3] = -0.25; prof[:, 4] = 0.08; prof[:, 7] = -0.01
profl = [list(r) for r in prof]
m = 6
xyz = np.hstack([tools.uv_to_xyz(s["pts"][:, t * 2 * m:(t + 1) * 2 * m], profl, s["dlt"]) for t in range(2)])
err = tools.get_repo_errors(xyz, s["pts"], profl, s["dlt"])
out.update(B_pts=s["pts"], B_prof=prof, B_dlt=s["dlt"], B_xyz=xyz, B_err=err)
# undistortion alone (float32 in/out), incl. far-off-image pixels
rng = np.random.default_rng(0)
px = np.concatenate([rng.uniform(-200, 2200, size=(3000, 2)), rng.uniform(0, 1920, size=(1000, 2))])
und = tools.undistort_pts(px, prof[0])
out.update(U_px=px, U_prof=prof[0], U_out=und)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "dlt_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
