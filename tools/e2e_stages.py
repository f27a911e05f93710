"""Where the time of a one-shot call goes (create / upload / solve / download / destroy through the resident API), per workload."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from argus_gui_b200 import ba
for wl in sys.argv[1:] or ["c4r8", "c4"]:
    s, p0, rot0, (m, n, views, nk, nd) = bench.make_scene(wl, 0)
    hp = torch.from_numpy(p0.copy()).pin_memory().numpy(); hx = torch.from_numpy(np.ascontiguousarray(s["x"])).pin_memory().numpy()
    hv = torch.from_numpy(np.ascontiguousarray(s["vmask"]).view(np.int8)).pin_memory().numpy()
    best = None
    for rep in range(2):
        t = [time.perf_counter()]
        B = ba.BundleAdjuster(device=0); t.append(time.perf_counter())
        B.upload(hv, hp, rot0, hx, nk, nd); t.append(time.perf_counter())
        B.solve(itmax=10, opts=bench.OPTS_FIXED_WORK); t.append(time.perf_counter())
        B.download(out=hp.copy()); t.append(time.perf_counter())
        B.close(); t.append(time.perf_counter())
        st = [1e3 * (b - a) for a, b in zip(t, t[1:])]
        best = st if best is None else [min(a, b) for a, b in zip(best, st)]
    print("%s (%d observations): create %.2f upload %.2f solve %.2f download %.2f close %.2f ms" % ((wl, s["nobs"]) + tuple(best)))
    one = []
    for rep in range(12):
        t0 = time.perf_counter()
        ba.solve_motstr(hv, hp, rot0, hx, nk, nd, itmax=10, opts=bench.OPTS_FIXED_WORK, inplace=True)
        one.append(1e3 * (time.perf_counter() - t0))
    print("   one-shot argus_ba_motstr_kdrts, 12 calls: " + " ".join("%.1f" % v for v in one) + " ms")
