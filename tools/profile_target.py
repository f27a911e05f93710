"""One upload + a 3-iteration solve of BASELINE config 4: the command profiled under ncu (after it has run plainly)."""
import sys; sys.path.insert(0, '/root/repo')
import bench
from argus_gui_b200 import ba
s, p0, rot0, (m, n, views, nk, nd) = bench.make_scene(sys.argv[1] if len(sys.argv) > 1 else "c4", 0)
B = ba.BundleAdjuster(device=0)
B.upload(s["vmask"], p0, rot0, s["x"], nk, nd)
print(B.solve(itmax=3, opts=bench.OPTS_FIXED_WORK))
B.close()
