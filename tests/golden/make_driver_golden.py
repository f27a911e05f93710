"""Generates tests/golden/driver_golden.npz by running the REFERENCE's sbaArgusDriver constructor (argus_gui/sbaDriver.py:
parse / count / rearrange / getPointsAndExtArray + triangulate.py) on a small wand scene.  GUI toolkits (PySide6,
pyqtgraph), texttable, `argus.ocam` and python-sba are not installed: they are stubbed (python-sba's quaternions by the
repo's re-implementation); none of them is executed by the constructor.  Run in the build container."""
import os, sys, types, importlib.util
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from argus_gui_b200 import scenes, sba
REF = os.environ.get("ARGUS_REFERENCE", "/root/reference")


def stub(name, **attrs):
    m = types.ModuleType(name); m.__dict__.update(attrs); sys.modules[name] = m; return m


class _Any(object):
    def __init__(self, *a, **k): pass
    def __getattr__(self, n): return _Any()
    def __call__(self, *a, **k): return _Any()


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
w = scenes.make_wand_scene(nframes=300, nbackground=40, seed=7)
out = {"ppts": w["ppts"], "uppts": w["uppts"], "cams": w["cams"]}
d = ref.sbaArgusDriver(w["ppts"].copy(), w["uppts"].copy(), w["cams"].copy(), display=False, scale=0.5, modeString="01", name="x", temp="", report=False)
out.update(pts=d.pts, ext=d.ext, order=d.order, nppts=d.nppts, nuppts=d.nuppts, cams_after=d.cams, ppts_after=d.ppts, uppts_after=d.uppts,
           paired0=np.array(d.indices["paired"][0]), paired1=np.array(d.indices["paired"][1]), unpaired0=np.array(d.indices["unpaired"][0]),
           counts=d.count(w["ppts"]) + d.count(w["uppts"]))
np.savez_compressed(os.path.join(HERE, "driver_golden.npz"), **out)
print({k: np.asarray(v).shape for k, v in out.items()}, d.order)
