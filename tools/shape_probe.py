"""Per-kernel times of one LM iteration for a few problem shapes (cameras, points, views)."""
import sys; sys.path.insert(0, '/root/repo')
import bench
from argus_gui_b200 import ba, scenes
for (m, n, v) in ((3, 1_000_000, 3), (4, 2_000_000, 4), (8, 2_000_000, 6), (16, 2_000_000, 10), (24, 1_000_000, 12)):
    s = scenes.make_ba_scene(m, n, v, seed=m)
    p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
    B = ba.BundleAdjuster().upload(s["vmask"], p0, rot0, s["x"], 0, 0)
    B.solve(itmax=3, opts=bench.OPTS_FIXED_WORK); B.reset()
    ret, info = B.solve(itmax=6, opts=bench.OPTS_FIXED_WORK, profile=True)
    pr = {k: round(v_[0] / max(v_[1], 1), 3) for k, v_ in B.profile().items() if v_[1]}
    tot = sum(v_[0] / v_[1] for k_, v_ in B.profile().items() if v_[1] and k_ not in ('residual', 'linearize_stats'))     # one attempt
    print("m=%d n=%d views=%d nobs=%d: %.2f ms/iter, %.3g obs*iter/s  %s" % (m, n, v, s["nobs"], tot, s["nobs"] / (tot * 1e-3), pr))
    B.close()
