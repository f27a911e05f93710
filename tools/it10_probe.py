import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from argus_gui_b200 import ba
s, p0, rot0, (m, n, views, nk, nd) = bench.make_scene("c4", 0)
B = ba.BundleAdjuster(device=0)
B.upload(s["vmask"], p0, rot0, s["x"], nk, nd)
for it in (5, 10, 20):
    B.reset(); t=time.perf_counter(); ret, info = B.solve(itmax=it, opts=bench.OPTS_FIXED_WORK); dt=time.perf_counter()-t
    print(it, ret, info, "%.1f ms  %.2f ms/iter" % (dt*1e3, dt*1e3/it))
B.reset(); t=time.perf_counter(); ret, info = B.solve(itmax=150); dt=time.perf_counter()-t
print("default opts:", ret, info, "%.1f ms" % (dt*1e3))
