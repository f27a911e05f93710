"""TEST INFRASTRUCTURE - numpy restatement of the reference's normal-equation algebra.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Restates, for given Jacobian blocks A_ij (2x16), B_ij (2x3) and residuals e_ij (the outputs of the reference's
img_projsKDRTS_jac_x / img_projsKDRTS_x, or of oracle/ba_port.c), what sba_motstr_levmar_x computes between the
Jacobian evaluation and the parameter update:

  U_j, ea_j, V_i, eb_i, W_ij          sba-1.6p/sba_levmar.c:837-938
  damping, (V*_i)^-1, Y_ij, S, E      sba-1.6p/sba_levmar.c:992-1185
  S da = E (Cholesky), db_i           sba-1.6p/sba_levmar.c:1201-1256, sba_lapack.c:373-479

Dense numpy, O(nobs * m) memory - small cases only.
"""
import numpy as np


def normal_equations(A, B, e, pt_idx, cam_idx, n, m, mu, mcon=0, ncon=0):
    A = np.asarray(A, dtype=np.float64).copy()
    B = np.asarray(B, dtype=np.float64).copy()
    e = np.asarray(e, dtype=np.float64)
    A[cam_idx < mcon] = 0.0          # "All A_ij with j<mcon are assumed to be zero" (sba_levmar.c:451-457)
    B[pt_idx < ncon] = 0.0
    U = np.zeros((m, 16, 16)); ea = np.zeros((m, 16)); V = np.zeros((n, 3, 3)); eb = np.zeros((n, 3))
    np.add.at(U, cam_idx, np.einsum("kri,krj->kij", A, A))
    np.add.at(ea, cam_idx, np.einsum("kri,kr->ki", A, e))
    np.add.at(V, pt_idx, np.einsum("kri,krj->kij", B, B))
    np.add.at(eb, pt_idx, np.einsum("kri,kr->ki", B, e))
    W = np.einsum("kri,krj->kij", A, B)                      # nobs x 16 x 3
    Vs = V + mu * np.eye(3)
    free_pt = np.arange(n) >= ncon
    Vinv = np.zeros_like(Vs)
    Vinv[free_pt] = np.linalg.inv(Vs[free_pt])
    Y = np.einsum("kij,kjl->kil", W, Vinv[pt_idx])
    mf = m - mcon
    N = 16 * mf
    S = np.zeros((N, N)); E = np.zeros(N)
    for j in range(mcon, m):
        a = j - mcon
        S[16 * a:16 * a + 16, 16 * a:16 * a + 16] = U[j] + mu * np.eye(16)
        E[16 * a:16 * a + 16] = ea[j]
    order = np.argsort(pt_idx, kind="stable")
    bounds = np.searchsorted(pt_idx[order], np.arange(n + 1))
    for i in range(ncon, n):
        ks = order[bounds[i]:bounds[i + 1]]
        for k1 in ks:
            j = cam_idx[k1]
            if j < mcon:
                continue
            a = j - mcon
            E[16 * a:16 * a + 16] -= Y[k1] @ eb[i]
            for k2 in ks:
                jj = cam_idx[k2]
                if jj < mcon:
                    continue
                b = jj - mcon
                S[16 * a:16 * a + 16, 16 * b:16 * b + 16] -= Y[k1] @ W[k2].T
    out = {"U": U, "ea": ea, "V": V, "eb": eb, "W": W, "S": S, "E": E}
    try:
        L = np.linalg.cholesky(S)
        da = np.linalg.solve(L.T, np.linalg.solve(L, E))
        da_full = np.zeros((m, 16)); da_full[mcon:] = da.reshape(mf, 16)
        t = eb.copy()
        np.subtract.at(t, pt_idx, np.einsum("kij,ki->kj", W, da_full[cam_idx]))
        db = np.einsum("nij,nj->ni", Vinv, t)
        db[:ncon] = 0.0
        out["dp"] = np.concatenate([da_full.ravel(), db.ravel()])
        out["solved"] = True
    except np.linalg.LinAlgError:
        out["solved"] = False
    return out
