"""Generates tests/golden/blockdiag_golden.npz: motion-only and structure-only LM runs of the REFERENCE's own compiled
C code (oracle/_ref: sba_mot_levmar_x / sba_str_levmar_x with the img_projsKDRT_* / img_projsKDS_* callbacks).
Run in the build container:  python tests/golden/make_blockdiag_golden.py
"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from argus_gui_b200 import scenes
from oracle import sba_ref

out = {}
# M1: 4 cameras, 300 points seen by 3 cameras each; motion-only from the perturbed cameras, points at their perturbed start
s = scenes.make_ba_scene(4, 300, 3, seed=11)
p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
m = 4
out.update(M1_vmask=s["vmask"], M1_x=s["x"], M1_p0=p0, M1_rot0=rot0)
for tag, (nk, nd, mcon) in {"a": (0, 0, 0), "b": (2, 3, 1), "c": (5, 4, 0)}.items():
    for it in (1, 3, 10):
        cams, info, ret = sba_ref.ref_mot(p0[:16 * m], p0[16 * m:], s["x"], s["vmask"], rot0, nk, nd, itmax=it, mcon=mcon)
        out["M1%s_cams_it%d" % (tag, it)] = cams; out["M1%s_info_it%d" % (tag, it)] = info
        print("M1" + tag, it, ret, info)
# M2: rejected steps: tiny initial damping and a far start
s2 = scenes.make_ba_scene(4, 200, 3, seed=23, pert_scale=40, outlier_frac=0.1)
q0, r0 = scenes.pack_ba_problem(s2["cams0"], s2["pts0"])
opts = np.array([1e-9, 1e-12, 1e-12, 1e-12, 0.0])
out.update(M2_vmask=s2["vmask"], M2_x=s2["x"], M2_p0=q0, M2_rot0=r0, M2_opts=opts)
for it in (2, 5):
    cams, info, ret = sba_ref.ref_mot(q0[:16 * m], q0[16 * m:], s2["x"], s2["vmask"], r0, 0, 0, itmax=it, opts=opts)
    out["M2_cams_it%d" % it] = cams; out["M2_info_it%d" % it] = info
    print("M2", it, ret, info)
# S1: structure-only.  The reference's KDS callbacks read the vector part of the FULL camera quaternion from the camera row
# (no rot0): start from zero local rotation, so that quaternion is rot0 itself (scalar part >= 0)
assert (rot0.reshape(m, 4)[:, 0] >= 0).all()
cams_full = p0[:16 * m].reshape(m, 16).copy(); cams_full[:, 10:13] = rot0.reshape(m, 4)[:, 1:4]
out["S1_cams_full"] = cams_full.ravel()
for tag, ncon in {"a": 0, "b": 17}.items():
    for it in (1, 2, 3):
        pts, info, ret = sba_ref.ref_str(cams_full.ravel(), p0[16 * m:], s["x"], s["vmask"], itmax=it, ncon=ncon)
        out["S1%s_pts_it%d" % (tag, it)] = pts; out["S1%s_info_it%d" % (tag, it)] = info
        print("S1" + tag, it, ret, info)
# S2: rejected steps
c2 = q0[:16 * m].reshape(m, 16).copy(); assert (r0.reshape(m, 4)[:, 0] >= 0).all(); c2[:, 10:13] = r0.reshape(m, 4)[:, 1:4]
out["S2_cams_full"] = c2.ravel()
for it in (2, 4):
    pts, info, ret = sba_ref.ref_str(c2.ravel(), q0[16 * m:], s2["x"], s2["vmask"], itmax=it, opts=opts)
    out["S2_pts_it%d" % it] = pts; out["S2_info_it%d" % it] = info
    print("S2", it, ret, info)
np.savez_compressed(os.path.join(HERE, "blockdiag_golden.npz"), **out)
print(sorted(out))
