"""Generates tests/golden/fit_golden.npz by importing the REFERENCE's own argus_gui/sbaDriver.py and running
OutlierWindow.getCoefficients (sbaDriver.py:1027-1165) on synthetic cameras (needs /root/reference).
Run in the build container:  python tests/golden/make_fit_golden.py
"""
import os, sys, types, importlib.util, warnings
import numpy as np

REF = os.environ.get("ARGUS_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from argus_gui_b200 import scenes



def stub(name, **attrs):
    m = types.ModuleType(name); m.__dict__.update(attrs); sys.modules[name] = m; return m


class _Any(object):
    def __init__(self, *a, **k): pass
    def __getattr__(self, n): return _Any()
    def __call__(self, *a, **k): return _Any()


# the reference's tools.solve_dlt does not run under numpy >= 2 (tools.py:267 assigns 1-element arrays into a float row); its
# twin OutlierWindow.getCoefficients (sbaDriver.py:1027-1165: same system, same lstsq, u-only NaN rule, plus the 3-sigma rule)
# does.  GUI toolkits / texttable / python-sba / argus.ocam are stubbed as in make_driver_golden.py; none is executed here.
from argus_gui_b200 import sba
pkg = types.ModuleType("argus_gui"); pkg.__path__ = [os.path.join(REF, "argus_gui")]; sys.modules["argus_gui"] = pkg
res = types.ModuleType("argus_gui.resources"); res.__path__ = [os.path.join(REF, "argus_gui", "resources")]; res.__file__ = os.path.join(REF, "argus_gui", "resources", "__init__.py")
res.__spec__ = importlib.util.spec_from_file_location("argus_gui.resources", res.__file__, submodule_search_locations=res.__path__)
sys.modules["argus_gui.resources"] = res
sys.modules["sba"] = sba; sys.modules["sba.quaternions"] = sba.quaternions
stub("argus", ocam=stub("argus.ocam", PointUndistorter=object, ocam_model=object))
qtw = stub("PySide6.QtWidgets", QWidget=object, QApplication=_Any); qtc = stub("PySide6.QtCore", Qt=_Any())
stub("PySide6", QtWidgets=qtw, QtCore=qtc)
gl = stub("pyqtgraph.opengl"); stub("pyqtgraph", opengl=gl)
stub("texttable", Texttable=_Any)
for name in ("tools", "output", "triangulate", "sbaDriver"):
    spec = importlib.util.spec_from_file_location("argus_gui." + name, os.path.join(REF, "argus_gui", name + ".py"))
    mod = importlib.util.module_from_spec(spec); sys.modules["argus_gui." + name] = mod; spec.loader.exec_module(mod)
ref = sys.modules["argus_gui.sbaDriver"]
fake = types.SimpleNamespace(ncams=3, order=None, indices={"paired": None, "unpaired": None}, nRef=0, nppts=0, nuppts=0, cams=None)

out = {}
for tag, (n, seed, nan_frac, noise) in {"A": (500, 3, 0.1, 0.5), "B": (60, 8, 0.0, 0.0), "C": (3000, 21, 0.3, 2.0)}.items():
    s = scenes.make_dlt_scene(n, m=3, ntracks=1, seed=seed, nan_frac=0.0)
    rng = np.random.default_rng(seed)
    xyz = s["xyz_true"].reshape(n, 3)
    L0 = s["dlt"]
    for j in range(3):                       # undistorted pixels of camera j = DLT projection of the true points + noise
        den = xyz @ L0[j, 8:11] + 1.0
        uv = np.stack([(xyz @ L0[j, 0:3] + L0[j, 3]) / den, (xyz @ L0[j, 4:7] + L0[j, 7]) / den], axis=1) + rng.normal(0, noise, (n, 2))
        uv[rng.random(n) < nan_frac] = np.nan
        import io, contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            L, rmse, outliers, ptsi = ref.OutlierWindow.getCoefficients(fake, xyz.copy(), uv.copy(), j)
        out.update({"%s%d_xyz" % (tag, j): xyz, "%s%d_uv" % (tag, j): uv, "%s%d_L" % (tag, j): np.asarray(L).ravel(), "%s%d_rmse" % (tag, j): rmse,
                    "%s%d_flagged" % (tag, j): np.asarray(ptsi, dtype=np.int64)})
        print(tag, j, rmse, np.asarray(L).ravel()[:4])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fit_golden.npz"), **out)
