"""Generates tests/golden/ba_golden.npz by running the REFERENCE's own compiled C code (oracle/_ref, built by
oracle/build_ref.sh from /root/reference/argus_gui/resources/sba_lib_sources.tar.gz) on small synthetic scenes.
Run in the build container:  python tests/golden/make_ba_golden.py
"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from argus_gui_b200 import scenes
from oracle import sba_ref

out = {}
# G1: 4 cameras, 120 points, 3 views; projection / Jacobian at a point with non-zero local rotations; LM after 1, 3, 8 iterations
s = scenes.make_ba_scene(4, 120, 3, seed=11)
p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
rng = np.random.default_rng(5)
p0r = p0.copy(); p0r[:64].reshape(4, 16)[:, 10:13] = rng.normal(0, 0.02, size=(4, 3))
out.update(G1_vmask=s["vmask"], G1_x=s["x"], G1_p0=p0, G1_p0r=p0r, G1_rot0=rot0)
out["G1_hx"] = sba_ref.ref_project(p0r, s["vmask"], rot0)
for nk, nd in ((0, 0), (2, 3), (5, 4)):
    A, B = sba_ref.ref_jacobian(p0r, s["vmask"], rot0, nk, nd)
    out["G1_A_%d%d" % (nk, nd)] = A; out["G1_B_%d%d" % (nk, nd)] = B
for it in (1, 3, 8):
    p, info, ret = sba_ref.ref_motstr(p0, s["x"], s["vmask"], rot0, 2, 3, itmax=it)
    out["G1_p_it%d" % it] = p; out["G1_info_it%d" % it] = info
# G2: rejected steps (tiny initial damping, far start, outliers): 6 iterations need 12 linear systems
s = scenes.make_ba_scene(4, 150, 3, seed=23, pert_scale=40, outlier_frac=0.1)
p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
opts = np.array([1e-8, 1e-12, 1e-12, 1e-12, 0.0])
out.update(G2_vmask=s["vmask"], G2_x=s["x"], G2_p0=p0, G2_rot0=rot0, G2_opts=opts)
for it in (2, 6):
    p, info, ret = sba_ref.ref_motstr(p0, s["x"], s["vmask"], rot0, 0, 0, itmax=it, opts=opts, mcon=1)
    out["G2_p_it%d" % it] = p; out["G2_info_it%d" % it] = info
    print("G2", it, info)
# G3: Argus wand mode '01' (intrinsics fixed, r2 free), 3 cameras, to convergence
s = scenes.make_ba_scene(3, 200, 3, seed=31, dist=(-0.05, 0, 0, 0, 0))
p0, rot0 = scenes.pack_ba_problem(s["cams0"], s["pts0"])
p, info, ret = sba_ref.ref_motstr(p0, s["x"], s["vmask"], rot0, 5, 4, itmax=150)
out.update(G3_vmask=s["vmask"], G3_x=s["x"], G3_p0=p0, G3_rot0=rot0, G3_p=p, G3_info=info)
print("G3", info)
np.savez_compressed(os.path.join(HERE, "ba_golden.npz"), **out)
print(sorted(out))
