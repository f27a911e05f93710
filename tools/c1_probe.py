"""BASELINE config 1 (3-camera wand calibration, ~4100 points, ~12000 observations): the bundle adjustment alone, GPU vs the
reference's C code on the same arrays (when oracle/_ref is present)."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np
from argus_gui_b200 import scenes, ba
s = scenes.make_ba_scene(3, 4100, 3, seed=1, dist=(-0.05, 0, 0, 0, 0))
p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
ba.solve_motstr(s["vmask"], p0, rot0, s["x"], 5, 4, itmax=150)
t = time.perf_counter(); p, info, ret = ba.solve_motstr(s["vmask"], p0, rot0, s["x"], 5, 4, itmax=150); dt = time.perf_counter() - t
print("gpu one-shot: %d iterations, %.2f ms total, %.3f ms/iteration, stop %d, cost %.6g" % (ret, 1e3 * dt, 1e3 * dt / max(ret, 1), info[6], info[1]))
B = ba.BundleAdjuster().upload(s["vmask"], p0, rot0, s["x"], 5, 4)
B.solve(itmax=150); B.reset()
t = time.perf_counter(); ret, info = B.solve(itmax=150); dt = time.perf_counter() - t
print("gpu resident: %d iterations, %.2f ms, %.3f ms/iteration" % (ret, 1e3 * dt, 1e3 * dt / max(ret, 1)))
import time as _t
st = {}
for rep in range(3):
    t0 = _t.perf_counter(); B2 = ba.BundleAdjuster(); B2.set_graph(False); t1 = _t.perf_counter()
    B2.upload(s["vmask"], p0, rot0, s["x"], 5, 4); t2 = _t.perf_counter()
    B2.solve(itmax=150); t3 = _t.perf_counter(); B2.download(); t4 = _t.perf_counter(); B2.close(); t5 = _t.perf_counter()
    st = {"create": t1 - t0, "upload": t2 - t1, "solve": t3 - t2, "download": t4 - t3, "destroy": t5 - t4}
print("stages of a one-shot call (ms):", {k: round(1e3 * v, 3) for k, v in st.items()})
try:
    from oracle import sba_ref
    if sba_ref.available():
        t = time.perf_counter(); pr, ir, rr = sba_ref.ref_motstr(p0, s["x"], s["vmask"], rot0, 5, 4, itmax=150); dtr = time.perf_counter() - t
        print("reference C: %d iterations, %.1f ms, cost %.6g" % (rr, 1e3 * dtr, ir[1]))
except Exception as e:
    print("reference unavailable:", e)
