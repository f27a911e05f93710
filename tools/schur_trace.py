"""Debug aid (needs a library built with -DARGUS_SCHUR_TRACE, see tools/build_trace.sh): per-warp clock stamps of the Schur kernel's
tile loop for the first CTAs -> who waits for whom at the stage barriers."""
import ctypes, sys
import numpy as np
sys.path.insert(0, '/root/repo')
import torch
import bench
from argus_gui_b200 import ba, _lib
wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
s, p0, rot0, (m, n, views, nk, nd) = bench.make_scene(wl, 0)
lib = _lib.load()
NW = 12
buf = torch.zeros(4 * 256 * NW * 4, dtype=torch.int64, device="cuda")
lib.argus_debug_set_schur_trace.argtypes = [ctypes.c_void_p]
B = ba.BundleAdjuster(device=0)
B.upload(s["vmask"], p0, rot0, s["x"], nk, nd)
B.set_graph(False)
B.solve(itmax=2, opts=bench.OPTS_FIXED_WORK)
assert lib.argus_debug_set_schur_trace(ctypes.c_void_p(buf.data_ptr())) == 0
B.reset(); B.solve(itmax=1, opts=bench.OPTS_FIXED_WORK)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(4, 256, NW, 4).astype(np.float64)
np.save("gpurun_out/schur_trace_%s.npy" % wl, t)
for cta in range(4):
    a = t[cta, 8:250]
    wait = a[:, :, 1] - a[:, :, 0]
    work = a[:, :, 2] - a[:, :, 1]
    per_tile = np.diff(a[:, 0, 2])
    print("cta %d: clk per tile %.0f; per warp work mean %s" % (cta, per_tile.mean(), np.round(work.mean(axis=0)).astype(int)))
    print("        per warp wait mean %s" % np.round(wait.mean(axis=0)).astype(int))
    # load latency: issue stamp of tile it+1 (slot 3, stored at index it+1 by warp 0) -> first warp passing the wait
    issue = a[1:, 0, 3]; ready = a[1:, :, 1].min(axis=1)
    print("        load issue -> first warp through: mean %.0f clk; issue relative to own tile start (warp 0): %.0f" % ((ready - issue).mean(), (issue - a[:-1, 0, 1]).mean()))
B.close()
