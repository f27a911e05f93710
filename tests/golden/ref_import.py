"""Imports the REFERENCE's own modules (argus_gui/tools.py, output.py, triangulate.py, sbaDriver.py) in the build container
for the golden generators.  What is not installed here is stubbed: the GUI toolkits (PySide6, pyqtgraph), texttable and
`argus.ocam` (none of them is executed on the paths the generators run) and python-sba, whose place is taken by this
repo's `argus_gui_b200.sba` containers with - when `backend="reference"` - the solve routed to the reference's own compiled C
code (oracle/_ref: sba_motstr_levmar_x + img_projsKDRTS_x / _jac_x), i.e. exactly what python-sba binds."""
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REF = os.environ.get("ARGUS_REFERENCE", "/root/reference")


class _Any(object):
    def __init__(self, *a, **k): pass
    def __getattr__(self, n): return _Any()
    def __call__(self, *a, **k): return _Any()


def _stub(name, **attrs):
    m = types.ModuleType(name); m.__dict__.update(attrs); sys.modules[name] = m; return m


def reference_backend_for_sba():
    """Route argus_gui_b200.sba.SparseBundleAdjust (motion + structure) to the reference's compiled C code."""
    from argus_gui_b200 import ba
    from oracle import sba_ref

    def solve_motstr(vmask, p, rot0, x, nccalib=0, ncdist=0, itmax=150, opts=None, ncon=0, mcon=0, verbose=0, compat=True, inplace=False):
        return sba_ref.ref_motstr(p, x, vmask, rot0, nccalib, ncdist, itmax=itmax, opts=list(opts) if opts is not None else None,
                                  ncon=ncon, mcon=mcon, verbose=verbose)
    ba.solve_motstr = solve_motstr


def load_reference(backend="reference"):
    """-> dict of the reference's modules {tools, output, triangulate, sbaDriver}."""
    from argus_gui_b200 import sba
    if backend == "reference":
        reference_backend_for_sba()
    pkg = types.ModuleType("argus_gui"); pkg.__path__ = [os.path.join(REF, "argus_gui")]; sys.modules["argus_gui"] = pkg
    res = types.ModuleType("argus_gui.resources"); res.__path__ = [os.path.join(REF, "argus_gui", "resources")]
    res.__file__ = os.path.join(REF, "argus_gui", "resources", "__init__.py")
    res.__spec__ = importlib.util.spec_from_file_location("argus_gui.resources", res.__file__, submodule_search_locations=res.__path__)
    sys.modules["argus_gui.resources"] = res
    sys.modules["sba"] = sba; sys.modules["sba.quaternions"] = sba.quaternions
    _stub("argus", ocam=_stub("argus.ocam", PointUndistorter=object, ocam_model=object))
    qtw = _stub("PySide6.QtWidgets", QWidget=object, QApplication=_Any()); qtc = _stub("PySide6.QtCore", Qt=_Any())
    _stub("PySide6", QtWidgets=qtw, QtCore=qtc)
    gl = _stub("pyqtgraph.opengl"); _stub("pyqtgraph", opengl=gl)
    _stub("texttable", Texttable=_Any)
    mods = {}
    cwd = os.getcwd()
    for name in ("tools", "output", "triangulate", "sbaDriver"):
        spec = importlib.util.spec_from_file_location("argus_gui." + name, os.path.join(REF, "argus_gui", name + ".py"))
        mod = importlib.util.module_from_spec(spec); sys.modules["argus_gui." + name] = mod; spec.loader.exec_module(mod)
        mods[name] = mod
    os.chdir(cwd)          # the reference's sbaDriver chdir()s into its resources directory on import (sbaDriver.py:10-31)
    return mods
