"""Debug aid (needs tools/_ab/liblintrace.so, built by tools/build_trace.sh lin): per-warp clock stamps at the phase boundaries of
the tile lineariser for the first tiles of two CTAs -> clocks per phase and per barrier wait."""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, '/root/repo')
from argus_gui_b200 import _lib
_lib.LIB_PATH = os.path.join('/root/repo', 'tools', '_ab', 'liblintrace.so')
import torch
import bench
from argus_gui_b200 import ba
wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
s, p0, rot0, (m, n, views, nk, nd) = bench.make_scene(wl, 0)
lib = _lib.load()
buf = torch.zeros(2 * 64 * 8 * 10, dtype=torch.int64, device="cuda")
lib.argus_debug_set_lin_trace.argtypes = [ctypes.c_void_p]
B = ba.BundleAdjuster(device=0)
B.upload(s["vmask"], p0, rot0, s["x"], nk, nd)
B.set_graph(False)
B.solve(itmax=2, opts=bench.OPTS_FIXED_WORK)
assert lib.argus_debug_set_lin_trace(ctypes.c_void_p(buf.data_ptr())) == 0
B.reset(); B.solve(itmax=1, opts=bench.OPTS_FIXED_WORK)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(2, 64, 8, 10).astype(np.float64)
names = ["load+bar", "P1 jacobian", "bar", "P2 points", "bar", "P3 Z", "bar", "P4 gram", "bar"]
for cta in range(2):
    a = t[cta, 4:60]                                   # (tile, warp, stamp)
    d = np.diff(a, axis=2)                             # per-phase clocks
    print("cta %d: clocks per tile %.0f" % (cta, np.diff(a[:, 0, 0]).mean()))
    for k, nm in enumerate(names):
        print("   %-12s mean over warps %7.0f   min warp %7.0f   max warp %7.0f" % (nm, d[:, :, k].mean(), d[:, :, k].mean(axis=0).min(), d[:, :, k].mean(axis=0).max()))
B.close()
