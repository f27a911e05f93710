"""c3 with the wand constraint on / off: ms per LM iteration and the per-kernel profile (device timed)."""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from argus_gui_b200 import ba, scenes
m, n, views, nk, nd = bench.WORKLOADS["c3"]
sw = scenes.make_ba_wand_scene(m, n // 2, 0, views, seed=4, cam_seed=4)
p0, rot0 = scenes.pack_ba_problem(sw["cams0"], sw["pts0"])
B = ba.BundleAdjuster(device=0).upload(sw["vmask"], p0, rot0, sw["x"], nk, nd)
for w in (0.0, 1000.0):
    B.set_wand(sw["ia"], sw["ib"], sw["wand"], w)
    for _ in range(2):
        B.reset(); B.solve(itmax=10, opts=bench.OPTS_FIXED_WORK)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        B.reset(); ret, info = B.solve(itmax=10, opts=bench.OPTS_FIXED_WORK)
    e1.record(); e1.synchronize()
    B.reset(); B.solve(itmax=10, opts=bench.OPTS_FIXED_WORK, profile=True)
    print("weight", w, "ms/iter %.3f" % (e0.elapsed_time(e1) / 30), {k: round(v[0] / v[1], 4) for k, v in B.profile().items() if v[1]})
B.close()
