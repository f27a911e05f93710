import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from argus_gui_b200 import ba
s, p0, rot0, (m, n, views, nk, nd) = bench.make_scene("c4", 0)
for _ in range(3):
    t = time.perf_counter()
    p, info, ret = ba.solve_motstr(s["vmask"], p0, rot0, s["x"], nk, nd, itmax=5, opts=bench.OPTS_FIXED_WORK)
    print("one-shot total %.1f ms" % (1e3 * (time.perf_counter() - t)), ret)
