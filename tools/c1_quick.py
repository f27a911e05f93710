"""Config 1 latency (3 cameras, wand-sized): resident solve and the one-shot call."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np
from argus_gui_b200 import scenes, ba
s = scenes.make_ba_scene(3, 4100, 3, seed=1, dist=(-0.05, 0, 0, 0, 0))
p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
B = ba.BundleAdjuster(device=0); B.upload(s["vmask"], p0, rot0, s["x"], 5, 4)
for _ in range(3): B.reset(); B.solve(itmax=150)
import torch; torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): B.reset(); ret, info = B.solve(itmax=150)
torch.cuda.synchronize(); ms = 1e3 * (time.perf_counter() - t0) / 20
B.reset(); ret, info = B.solve(itmax=150, profile=True); prof = B.profile(); B.close()
one = []
for _ in range(5):
    t0 = time.perf_counter(); ba.solve_motstr(s["vmask"], p0, rot0, s["x"], 5, 4, itmax=150); one.append(1e3 * (time.perf_counter() - t0))
print("c1: %d iterations, resident %.3f ms (%.4f ms / iteration), one-shot call %.3f ms" % (ret, ms, ms / ret, min(one)))
print({k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items() if v[1]})
