"""TEST INFRASTRUCTURE - numpy restatement of the reference's DLT path (CPU oracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Restates, vectorised over frames but with the same arithmetic per frame:

  undistort_pts_f32   /root/reference/argus_gui/tools.py:129-148  ->  cv2.undistortPoints(src float32, K, dist, P=K):
                      OpenCV's 5 fixed-point iterations (default criteria MAX_ITER 5) in double, float32 in/out.
                      (third-party: opencv-contrib-python >= 4.0, unpinned by the reference; 4.13.0 in this image.)
  uv_to_xyz           /root/reference/argus_gui/tools.py:296-337  incl. the row-overwrite at :317-328: the system
                      solved per frame is the u-equation of every visible camera plus the v-equation of the LAST
                      visible camera; rows beyond n+1 stay zero.  numpy lstsq (SVD, min-norm) per frame.
  reconstruct_uv      /root/reference/argus_gui/tools.py:64-76
  get_repo_errors     /root/reference/argus_gui/tools.py:360-423
  bootstrap_sd        /root/reference/argus_gui/tools.py:513-530  (the perturb / re-triangulate / nanstd loop of bootstrapXYZs,
                      with the standard-normal numbers passed in explicitly instead of drawn from np.random)

Pinned against the reference itself: tests/golden/dlt_golden.npz was produced by importing the reference's
tools.py (tests/golden/make_dlt_golden.py) and tests/test_oracle_dlt.py checks this file against it.
"""
import numpy as np


def undistort_pts_f32(pts, prof):
    """pts (N,2) float64 pixels -> (N,2) float32, as tools.undistort_pts for a pinhole profile
    prof = [f, cx, cy, k1, k2, p1, p2, k3] (only prof[0:3] and prof[-5:] are used, tools.py:134-140)."""
    prof = np.asarray(prof, dtype=np.float64)
    fx, cx, cy = prof[0], prof[1], prof[2]
    k1, k2, p1, p2, k3 = prof[-5:]
    src = np.asarray(pts, dtype=np.float64).astype(np.float32).astype(np.float64)
    u, v = src[:, 0], src[:, 1]
    ifx = 1.0 / fx
    x = (u - cx) * ifx
    y = (v - cy) * ifx
    x0, y0 = x.copy(), y.copy()
    done = np.zeros(x.shape, dtype=bool)
    with np.errstate(all="ignore"):
        for _ in range(5):
            r2 = x * x + y * y
            icdist = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2)
            neg = (icdist < 0) & ~done
            dX = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x)
            dY = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y
            xn = (x0 - dX) * icdist
            yn = (y0 - dY) * icdist
            xn = np.where(neg, x0, xn)
            yn = np.where(neg, y0, yn)
            x = np.where(done, x, xn)
            y = np.where(done, y, yn)
            done |= neg
        xx = fx * x + 0.0 * y + cx
        yy = 0.0 * x + fx * y + cy
    return np.stack([xx, yy], axis=1).astype(np.float32)


def uv_to_xyz(pts, profs, dlt, textbook=False):
    """pts (N, 2m) -> xyz (N,3); per-frame lstsq exactly as the reference (or the textbook 2n-row system)."""
    pts = np.asarray(pts, dtype=np.float64)
    dlt = np.asarray(dlt, dtype=np.float64)
    N, m = pts.shape[0], pts.shape[1] // 2
    und = np.stack([undistort_pts_f32(pts[:, 2 * j:2 * j + 2], profs[j]).astype(np.float64) for j in range(m)], axis=1)  # N,m,2
    vis = ~np.isnan(pts.reshape(N, m, 2)).any(axis=2)
    xyzs = np.zeros((N, 3))
    for i in range(N):
        cams = np.nonzero(vis[i])[0]
        n = len(cams)
        if n > 1:
            A = np.zeros((2 * n, 3)); B = np.zeros((2 * n, 1))
            if textbook:
                for k, j in enumerate(cams):
                    L = dlt[j]; uu, vv = und[i, j]
                    A[2 * k] = uu * L[8:11] - L[0:3]; B[2 * k] = L[3] - uu
                    A[2 * k + 1] = vv * L[8:11] - L[4:7]; B[2 * k + 1] = L[7] - vv
            else:
                for k, j in enumerate(cams):     # rows k and k+1: the next k overwrites row k+1
                    L = dlt[j]; uu, vv = und[i, j]
                    A[k] = uu * L[8:11] - L[0:3]; B[k] = L[3] - uu
                    A[k + 1] = vv * L[8:11] - L[4:7]; B[k + 1] = L[7] - vv
            xyzs[i] = np.linalg.lstsq(A, B, rcond=None)[0][:, 0]
    xyzs[xyzs == 0] = np.nan
    return xyzs


def reconstruct_uv(L, xyz):
    den = np.dot(L[-3:], xyz) + 1.0
    return np.array([(np.dot(L[:3], xyz) + L[3]) / den, (np.dot(L[4:7], xyz) + L[7]) / den])


def get_repo_errors(xyzs, pts, prof, dlt):
    """xyzs (N,3k), pts (N, 2mk) -> (k, N)."""
    xyzs = np.asarray(xyzs, dtype=np.float64); pts = np.asarray(pts, dtype=np.float64); dlt = np.asarray(dlt, dtype=np.float64)
    m = len(prof)
    k = xyzs.shape[1] // 3
    N = xyzs.shape[0]
    out = np.zeros((k, N))
    with np.errstate(all="ignore"):
        for t in range(k):
            xyz = xyzs[:, 3 * t:3 * t + 3]
            uv = pts[:, t * 2 * m:(t + 1) * 2 * m]
            und = [undistort_pts_f32(uv[:, 2 * j:2 * j + 2], prof[j]).astype(np.float64) for j in range(m)]
            for i in range(N):
                if np.isnan(xyz[i]).any():
                    continue
                eps = 0.0; n = 0
                for j in range(m):
                    if not np.isnan(uv[i, 2 * j]):
                        re = reconstruct_uv(dlt[j], xyz[i])
                        eps += (und[j][i, 0] - re[0]) ** 2 + (und[j][i, 1] - re[1]) ** 2
                        n += 1
                if n == 2:
                    out[t, i] = np.sqrt(eps / 2.0)
                else:
                    out[t, i] = np.sqrt(np.float64(eps) / float(2 * n - 3))
    out[out == 0] = np.nan
    return out


def bootstrap_sd(pts, rmses, prof, dlt, noise):
    """pts N x 2mk, rmses N x k, noise (k, bsIter, N, 2m) -> N x 3k standard deviations (ddof = 1, NaN-aware), exactly the
    loop of bootstrapXYZs (tools.py:513-530) for given random numbers."""
    import warnings
    pts = np.asarray(pts, dtype=np.float64)
    m = len(prof)
    N = pts.shape[0]
    k = pts.shape[1] // (2 * m)
    bs = noise.shape[1]
    ret = np.zeros((N, 3 * k))
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        for t in range(k):
            track = pts[:, t * 2 * m:(t + 1) * 2 * m]
            camSum = np.nansum(np.isfinite(track), axis=1) / 2
            xyzBS = np.full((N, 3, bs), np.nan)
            for j in range(bs):
                per = noise[t, j] * np.tile((rmses[:, t] * 2 ** 0.5 / camSum).reshape((-1, 1)), (1, 2 * m)) + track
                xyzBS[:, :, j] = uv_to_xyz(per, prof, dlt)
            sd = np.nanstd(xyzBS, axis=2, ddof=1)
            sd[sd == 0] = np.nan
            ret[:, t * 3:(t + 1) * 3] = sd
    return ret


def solve_dlt(xyz, uv, both_coordinates=True):
    """tools.solve_dlt (tools.py:220-280; rows with a NaN in uv dropped) / OutlierWindow.getCoefficients
    (sbaDriver.py:1027-1075; rows with a NaN in u dropped): numpy lstsq on the 2n x 11 system, reprojection distances.
    -> (L (11,), errors over the used rows, used-row indices)."""
    xyz = np.asarray(xyz, dtype=np.float64); uv = np.asarray(uv, dtype=np.float64)
    keep = ~np.isnan(uv).any(axis=1) if both_coordinates else ~np.isnan(uv[:, 0])
    idx = np.nonzero(keep)[0]
    X, w = xyz[idx], uv[idx]
    n = len(idx)
    A = np.zeros((2 * n, 11)); B = np.zeros((2 * n, 1))
    A[0::2, 0:3] = X; A[0::2, 3] = 1.0; A[0::2, 8:11] = -X * w[:, 0:1]
    A[1::2, 4:7] = X; A[1::2, 7] = 1.0; A[1::2, 8:11] = -X * w[:, 1:2]
    B[0::2, 0] = w[:, 0]; B[1::2, 0] = w[:, 1]
    L = np.linalg.lstsq(A, B, rcond=None)[0][:, 0]
    den = X @ L[8:11] + 1.0
    rec = np.stack([(X @ L[0:3] + L[3]) / den, (X @ L[4:7] + L[7]) / den], axis=1)
    return L, np.sqrt(((rec - w) ** 2).sum(axis=1)), idx
