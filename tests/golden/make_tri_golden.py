"""Generates tests/golden/tri_golden.npz by importing the REFERENCE's argus_gui/triangulate.py (needs /root/reference and
cv2).  The external python-sba module it imports only for ``sba.quaternions.quatFromRotationMatrix`` is stubbed with the
repo's re-implementation (argus_gui_b200/sba/quaternions.py); `argus.ocam` (omnidirectional model, unused) is stubbed.
Run in the build container:  python tests/golden/make_tri_golden.py"""
import os, sys, types, importlib.util
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from argus_gui_b200 import scenes, sba
REF = os.environ.get("ARGUS_REFERENCE", "/root/reference")
pkg = types.ModuleType("argus_gui"); pkg.__path__ = [os.path.join(REF, "argus_gui")]; sys.modules["argus_gui"] = pkg
sys.modules["sba"] = sba; sys.modules["sba.quaternions"] = sba.quaternions
argus = types.ModuleType("argus"); ocam = types.ModuleType("argus.ocam"); ocam.PointUndistorter = object; ocam.ocam_model = object
argus.ocam = ocam; sys.modules["argus"] = argus; sys.modules["argus.ocam"] = ocam
for name in ("tools", "triangulate"):
    spec = importlib.util.spec_from_file_location("argus_gui." + name, os.path.join(REF, "argus_gui", name + ".py"))
    mod = importlib.util.module_from_spec(spec); sys.modules["argus_gui." + name] = mod; spec.loader.exec_module(mod)
ref = sys.modules["argus_gui.triangulate"]
out = {}
for tag, (m, n, seed) in {"A": (3, 300, 5), "B": (4, 200, 9)}.items():
    s = scenes.make_ba_scene(m, n, m, seed=seed, radius=5.0)
    dense = np.full((n, 2 * m), np.nan); uv = dense.reshape(n, m, 2); uv[s["vmask"] != 0] = s["x"]
    rng = np.random.default_rng(seed); drop = rng.random((n, m)) < 0.1; drop[:, -1] = False; uv[drop] = np.nan
    intr = s["cams_true"][:, :10].copy()
    t = ref.multiTriangulator(dense.copy(), intr); xyz = t.triangulate()
    out.update({tag + "_pts": dense, tag + "_intr": intr, tag + "_xyz": xyz, tag + "_ext": t.ext})
np.savez_compressed(os.path.join(HERE, "tri_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
