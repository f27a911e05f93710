"""Generates tests/golden/bootstrap_golden.npz by importing the REFERENCE's argus_gui/tools.py and running bootstrapXYZs with a
seeded numpy generator; the standard-normal numbers it draws (np.random.randn, one (N, 2m) block per track and iteration) are
re-generated with the same seed and stored so that the CUDA path and the oracle can be fed the identical noise.
Run in the build container:  python tests/golden/make_bootstrap_golden.py"""
import os, sys, types, importlib.util
import numpy as np
REF = os.environ.get("ARGUS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from argus_gui_b200 import scenes
argus = types.ModuleType("argus"); ocam = types.ModuleType("argus.ocam")
ocam.PointUndistorter = type("PointUndistorter", (), {}); ocam.ocam_model = type("ocam_model", (), {}); argus.ocam = ocam
sys.modules["argus"] = argus; sys.modules["argus.ocam"] = ocam
spec = importlib.util.spec_from_file_location("ref_tools", os.path.join(REF, "argus_gui", "tools.py"))
tools = importlib.util.module_from_spec(spec); spec.loader.exec_module(tools)

N, m, k, bs = 60, 4, 2, 40
s = scenes.make_dlt_scene(N, m=m, ntracks=k, seed=11, nan_frac=0.15)
pts = s["pts"].copy(); pts[0, :2 * m] = np.nan; pts[1, 2:2 * m] = np.nan; pts[2, 1] = np.nan
prof = [list(r) for r in s["prof"]]
xyz = np.hstack([tools.uv_to_xyz(pts[:, t * 2 * m:(t + 1) * 2 * m], prof, s["dlt"]) for t in range(k)])
rmses = tools.get_repo_errors(xyz, pts, prof, s["dlt"]).T
np.random.seed(2024)
ci, w, tol = tools.bootstrapXYZs(pts.copy(), rmses, prof, s["dlt"], bsIter=bs, subframeinterp=False)
np.random.seed(2024)
noise = np.stack([np.stack([np.random.randn(N, 2 * m) for _ in range(bs)]) for _ in range(k)])
# sub-frame interpolation path (3+ cameras): camera 3's pixels are shifted by 0.3 frames on a smooth trajectory
N2 = 80
tt = np.linspace(0, 1, N2)
X = np.stack([0.8 * np.sin(2 * np.pi * tt), 0.8 * np.cos(2 * np.pi * tt), 0.5 * tt], axis=1)
cams = s["cams"]
pts2 = np.hstack([scenes.project_points(cams, X, np.full(N2, j)) for j in range(m)])
shift = np.stack([np.interp(np.arange(N2) - 0.3, np.arange(N2), pts2[:, 4 + q]) for q in range(2)], axis=1)
pts2[:, 4:6] = shift
xyz2 = tools.uv_to_xyz(pts2, prof, s["dlt"]); rm2 = tools.get_repo_errors(xyz2, pts2, prof, s["dlt"]).T
pts2_in = pts2.copy()
np.random.seed(7)
ci2, w2, tol2 = tools.bootstrapXYZs(pts2, rm2, prof, s["dlt"], bsIter=10, subframeinterp=True)
np.random.seed(7)
noise2 = np.stack([np.stack([np.random.randn(N2, 2 * m) for _ in range(10)])])
np.savez_compressed(os.path.join(HERE, "bootstrap_golden.npz"), pts=pts, prof=s["prof"], dlt=s["dlt"], rmses=rmses, noise=noise, ci=ci, w=w, tol=tol,
                    pts2_in=pts2_in, pts2_out=pts2, rm2=rm2, noise2=noise2, ci2=ci2, w2=w2, tol2=tol2)
print(ci.shape, np.isnan(ci).sum(), ci2.shape, np.abs(pts2 - pts2_in).max())
