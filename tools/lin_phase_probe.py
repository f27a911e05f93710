import sys, os; sys.path.insert(0, '/root/repo')
os.environ["ARGUS_LIN_DEBUG"] = "1"
import bench
from argus_gui_b200 import ba
s, p0, rot0, (m, n, views, nk, nd) = bench.make_scene("c4", 0)
B = ba.BundleAdjuster(device=0)
B.upload(s["vmask"], p0, rot0, s["x"], nk, nd)
ret, info = B.solve(itmax=10, opts=bench.OPTS_FIXED_WORK, profile=True)
print(ret, {k: (round(v[0] / max(v[1], 1), 3), v[1]) for k, v in B.profile().items()})
B.close()
