"""Small end-to-end run of every entry point, a quick health check of the built library on a GPU box."""
import sys; sys.path.insert(0, '/root/repo')
import numpy as np
from argus_gui_b200 import scenes, ba, dlt, triangulate
s = scenes.make_ba_scene(5, 700, 4, seed=3)
p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
for mode in ("motstr", "mot", "str"):
    B = ba.BundleAdjuster(mode=mode).upload(s["vmask"], p0, rot0, s["x"], 2, 3)
    print(mode, B.solve(itmax=3)); B.close()
s2 = scenes.make_ba_scene(33, 400, 12, seed=4)
q0, r0 = scenes.pack_ba_problem(s2["cams0"], s2["pts0"])
print("m=33", ba.solve_motstr(s2["vmask"], q0, r0, s2["x"], 0, 0, itmax=2)[2])
d = scenes.make_dlt_scene(3000, m=4, ntracks=3, seed=2)
xyz, err = dlt.reconstruct(d["pts"], d["prof"], d["dlt"])
x1 = dlt.triangulate(d["pts"], d["prof"], d["dlt"]); e1 = dlt.reproj_error(x1, d["pts"], d["prof"], d["dlt"])
print("dlt", np.nanmean(err), np.array_equal(x1, xyz, equal_nan=True))
rm = np.nan_to_num(err.T, nan=1.0)
print("boot", np.nanmean(dlt.bootstrap_sd(d["pts"][:500], rm[:500], d["prof"], d["dlt"], bs_iter=20, seed=1)))
L, e, st = dlt.fit(d["xyz_true"].reshape(3000, 9)[:, :3], np.nan_to_num(d["pts"][:, :2], nan=100.0))
print("fit", st[0])
sc = scenes.make_ba_scene(2, 3000, 2, seed=5, radius=5.0)
dense = sc["x"].reshape(3000, 4); intr = sc["cams_true"][:, :10]
T = triangulate.Triangulator(dense[:, :2], dense[:, 2:], intr[0, 0], intr[1, 0], intr[0, 1:3], intr[1, 1:3], intr[0, -5:], intr[1, -5:])
print("pair", np.isfinite(T.triangulate()).all())
