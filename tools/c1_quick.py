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
# where the one-shot call's time goes: the same steps through the resident API
def stage():
    t = [time.perf_counter()]
    B = ba.BundleAdjuster(device=0); t.append(time.perf_counter())
    B.upload(s["vmask"], p0, rot0, s["x"], 5, 4); t.append(time.perf_counter())
    B.solve(itmax=150); t.append(time.perf_counter())
    B.download(); t.append(time.perf_counter())
    B.close(); t.append(time.perf_counter())
    return [1e3 * (b - a) for a, b in zip(t, t[1:])]
best = None
for _ in range(5):
    st = stage(); best = st if best is None else [min(a, b) for a, b in zip(best, st)]
print("stages ms: create %.3f upload %.3f solve %.3f download %.3f close %.3f" % tuple(best))
def stage2():
    B = ba.BundleAdjuster(device=0); B.set_graph(False); B.upload(s["vmask"], p0, rot0, s["x"], 5, 4)
    t0 = time.perf_counter(); B.solve(itmax=150); t1 = time.perf_counter(); B.close(); return 1e3 * (t1 - t0)
print("solve without graph capture: %.3f ms" % min(stage2() for _ in range(5)))
