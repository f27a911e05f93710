"""Ad-hoc first-light check on a B200: product C-ABI vs the compiled reference (oracle/_ref)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argus_gui_b200 import scenes, ba
from oracle import sba_ref, ba_numpy

def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

for (m, n, views, nk, nd, mcon) in [(4, 500, 3, 2, 3, 0), (8, 3000, 6, 0, 0, 0), (3, 300, 3, 5, 4, 0), (6, 1000, 4, 2, 0, 1)]:
    s = scenes.make_ba_scene(m, n, views, seed=m * 7 + n)
    p0, rot0 = scenes.pack_ba_problem(s['cams0'], s['pts0'])
    rng = np.random.default_rng(1)
    p0r = p0.copy(); p0r[:16 * m].reshape(m, 16)[:, 10:13] = rng.normal(0, 0.01, size=(m, 3))
    B = ba.BundleAdjuster().upload(s['vmask'], p0r, rot0, s['x'], nk, nd, mcon=mcon)
    hx = B.project(); rhx = sba_ref.ref_project(p0r, s['vmask'], rot0)
    print("case", (m, n, views, nk, nd, mcon), "proj rel", rel(hx, rhx))
    rA, rB = sba_ref.ref_jacobian(p0r, s['vmask'], rot0, nk, nd)
    pi, ci = np.nonzero(s['vmask'])
    e = s['x'] - rhx
    mu = 0.37
    ne = ba_numpy.normal_equations(rA, rB, e, pi, ci, n, m, mu, mcon=mcon)
    g = B.normal_equations(mu, mcon=mcon)
    V6 = np.stack([ne['V'][:, 0, 0], ne['V'][:, 0, 1], ne['V'][:, 0, 2], ne['V'][:, 1, 1], ne['V'][:, 1, 2], ne['V'][:, 2, 2]], axis=1)
    print("  U", rel(g['U'], ne['U']), "ea", rel(g['ea'], ne['ea']), "V", rel(g['V'], V6), "eb", rel(g['eb'], ne['eb']),
          "S", rel(g['S'], ne['S']), "E", rel(g['E'], ne['E']), "solved", g['scal'][7], ne['solved'])
    if ne['solved']:
        print("  dp", rel(g['dp'], ne['dp']), "da", rel(g['dp'][:16*m], ne['dp'][:16*m]), "Ssym", np.abs(g['S'] - g['S'].T).max())
    # full LM trajectories
    p0, rot0 = scenes.pack_ba_problem(s['cams0'], s['pts0'])
    for itmax in (1, 3, 10, 60):
        t = time.time(); pr, ir, rr = sba_ref.ref_motstr(p0, s['x'], s['vmask'], rot0, nk, nd, itmax=itmax, mcon=mcon); tr = time.time() - t
        t = time.time(); pg, ig, rg = ba.solve_motstr(s['vmask'], p0, rot0, s['x'], nk, nd, itmax=itmax, mcon=mcon); tg = time.time() - t
        print("  itmax", itmax, "ret", rr, rg, "p rel", rel(pg, pr), "cost", ir[1], ig[1], "info5-9", ir[5:], ig[5:], "t ref %.3f gpu %.3f" % (tr, tg))
    B.close()
