"""TEST INFRASTRUCTURE - dense numpy model of the in-solver wand-length constraint.  PARITY UNPINNED.

Only tests/ may import this.  The reference has no counterpart (it applies the wand length after the bundle adjustment, as
one global scale: /root/reference/argus_gui/sbaDriver.py:703-713, and reports the spread as the wand score, :779-781), so
there is nothing to pin this model against; what it gives the tests is an independent statement of the mathematics: the FULL
Jacobian of [reprojection residuals ; wand residuals] as one dense matrix, and one Levenberg-Marquardt step from the dense
augmented normal equations (J^T J + mu I) dp = J^T e - no Schur complement, no Sherman-Morrison, none of the structure the CUDA
path exploits (argus_gui_b200/csrc/ba_wand.cuh).

Residual convention of sba (sba_levmar.c:791-1357): e = measurement - prediction, J = d prediction / d p.  A wand pair
(a, b) of length L and weight w: measurement w L, prediction w ||X_a - X_b||.
"""
import numpy as np


def wand_residuals(pts, ia, ib, length, w):
    d = np.linalg.norm(pts[ia] - pts[ib], axis=1)
    return w * (np.broadcast_to(length, d.shape) - d)


def wand_jacobian_rows(pts, ia, ib, w):
    """-> (npairs x 3) g_a; the row of pair p is +g_a[p] in the columns of point ia[p] and -g_a[p] in those of ib[p]."""
    diff = pts[ia] - pts[ib]
    return w * diff / np.linalg.norm(diff, axis=1, keepdims=True)


def dense_step(A, B, e, pt_idx, cam_idx, n, m, mu, pts, ia, ib, length, w, mcon=0, ncon=0):
    """One LM step of the constrained problem from dense matrices.  A (nobs x 2 x 16), B (nobs x 2 x 3), e (nobs x 2) are
    the reprojection Jacobian blocks and residuals (of the oracle's Jacobian, oracle/ba_port.py).  Returns the full Jacobian,
    residual vector, J^T e, the diagonal of J^T J, dp = [16 m | 3 n] (zeros for fixed cameras / points) and the cost."""
    nobs, npairs = A.shape[0], len(ia)
    ncol = 16 * m + 3 * n
    J = np.zeros((2 * nobs + npairs, ncol))
    r = np.concatenate([np.asarray(e).ravel(), wand_residuals(pts, ia, ib, length, w)])
    for k in range(nobs):
        j, i = cam_idx[k], pt_idx[k]
        J[2 * k:2 * k + 2, 16 * j:16 * j + 16] = A[k]
        J[2 * k:2 * k + 2, 16 * m + 3 * i:16 * m + 3 * i + 3] = B[k]
    g = wand_jacobian_rows(pts, ia, ib, w)
    for p in range(npairs):
        J[2 * nobs + p, 16 * m + 3 * ia[p]:16 * m + 3 * ia[p] + 3] = g[p]
        J[2 * nobs + p, 16 * m + 3 * ib[p]:16 * m + 3 * ib[p] + 3] = -g[p]
    free = np.ones(ncol, dtype=bool)
    free[:16 * mcon] = False
    free[16 * m:16 * m + 3 * ncon] = False
    Jf = J[:, free]
    H = Jf.T @ Jf
    gvec = Jf.T @ r
    dpf = np.linalg.solve(H + mu * np.eye(H.shape[0]), gvec)
    dp = np.zeros(ncol); dp[free] = dpf
    return {"J": J, "r": r, "free": free, "JTe": gvec, "diag": np.diag(H), "dp": dp, "cost": float(r @ r),
            "dL": float(dpf @ (mu * dpf + gvec))}
