"""Times the REFERENCE's own tools.uv_to_xyz + get_repo_errors (/root/reference/argus_gui/tools.py:296-423) beside the numpy
restatement bench.py uses as the DLT cpu_baseline (the reference tree does not travel to the GPU box, so bench.py cannot time
it there).  Run in the build container:  python tools/dlt_reference_timing.py"""
import os, sys, types, time, importlib.util
import numpy as np
REF = os.environ.get("ARGUS_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from argus_gui_b200 import scenes
from oracle import dlt_port
argus = types.ModuleType("argus"); ocam = types.ModuleType("argus.ocam")
ocam.PointUndistorter = type("PointUndistorter", (), {}); ocam.ocam_model = type("ocam_model", (), {})
argus.ocam = ocam
sys.modules["argus"] = argus; sys.modules["argus.ocam"] = ocam
spec = importlib.util.spec_from_file_location("ref_tools", os.path.join(REF, "argus_gui", "tools.py"))
tools = importlib.util.module_from_spec(spec); spec.loader.exec_module(tools)
N, m, k = 1500, 4, 10
d = scenes.make_dlt_scene(N, m=m, ntracks=k, seed=2)
prof = [list(r) for r in d["prof"]]
t0 = time.perf_counter()
xr = np.hstack([tools.uv_to_xyz(d["pts"][:, t * 2 * m:(t + 1) * 2 * m], prof, d["dlt"]) for t in range(k)])
er = tools.get_repo_errors(xr, d["pts"], prof, d["dlt"])
tr = time.perf_counter() - t0
t0 = time.perf_counter()
xp = np.hstack([dlt_port.uv_to_xyz(d["pts"][:, t * 2 * m:(t + 1) * 2 * m], d["prof"], d["dlt"]) for t in range(k)])
ep = dlt_port.get_repo_errors(xp, d["pts"], d["prof"], d["dlt"])
tp = time.perf_counter() - t0
print("reference tools.py: %.0f points/s (%.2f s);  numpy restatement: %.0f points/s (%.2f s);  identical xyz: %s, errors: %s"
      % (N * k / tr, tr, N * k / tp, tp, np.array_equal(xr, xp, equal_nan=True), np.array_equal(er, ep, equal_nan=True)))
