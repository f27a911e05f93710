import sys, os; sys.path.insert(0, '/root/repo')
import bench
from argus_gui_b200 import ba, scenes
for (m, n, v) in ((8, 100_000, 6), (16, 100_000, 10), (24, 100_000, 12), (32, 100_000, 12), (54, 100_000, 12), (64, 100_000, 12)):
    s = scenes.make_ba_scene(m, n, v, seed=m)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    B = ba.BundleAdjuster().upload(s["vmask"], p0, rot0, s["x"], 0, 0)
    B.solve(itmax=2, opts=bench.OPTS_FIXED_WORK); B.reset()
    ret, info = B.solve(itmax=4, opts=bench.OPTS_FIXED_WORK, profile=True)
    pr = B.profile()
    print("m=%d N=%d: chol %.3f ms, iteration %.3f ms" % (m, 16 * m, pr["chol_solve"][0] / pr["chol_solve"][1], sum(v_[0] / v_[1] for k_, v_ in pr.items() if v_[1] and k_ not in ('residual', 'linearize_stats'))))
    B.close()
