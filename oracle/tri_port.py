"""TEST INFRASTRUCTURE - CPU restatement (numpy + OpenCV, the reference's own dependencies) of the two-view initialisation:
/root/reference/argus_gui/triangulate.py:27-307 (Triangulator, multiTriangulator) with the per-point Python loops vectorised.
Pinned by tests/golden/tri_golden.npz (outputs of the reference module itself, tests/golden/make_tri_golden.py).
Only tests/ may import this module; the product path (argus_gui_b200/triangulate.py) calls the CUDA library.
"""
import numpy as np
import cv2

from argus_gui_b200.sba import quaternions


class ArgusError(Exception):
    pass

__all__ = ["gs", "Triangulator", "multiTriangulator"]


def gs(X, row_vecs=False, norm=True):
    """Gram-Schmidt orthonormalisation of the columns (rows if row_vecs) of X (triangulate.py:12-24)."""
    Q = np.array(X, dtype=np.float64, copy=True)
    if not row_vecs:
        Q = Q.T
    out = np.zeros_like(Q)
    for i in range(Q.shape[0]):
        v = Q[i].copy()
        for j in range(i):
            v -= (Q[i] @ out[j]) / (out[j] @ out[j]) * out[j]
        out[i] = v
    if norm:
        out = out / np.linalg.norm(out, axis=1, keepdims=True)
    return out if row_vecs else out.T


def _triangulate_pairs(P_a, P_b, xa, xb):
    """Batched two-view linear triangulation (what cv2.triangulatePoints does per point): for every point the null
    vector of the 4x4 system [x P3 - P1; y P3 - P2] for both views, de-homogenised."""
    n = xa.shape[0]
    A = np.empty((n, 4, 4))
    A[:, 0] = xa[:, 0:1] * P_a[2] - P_a[0]
    A[:, 1] = xa[:, 1:2] * P_a[2] - P_a[1]
    A[:, 2] = xb[:, 0:1] * P_b[2] - P_b[0]
    A[:, 3] = xb[:, 1:2] * P_b[2] - P_b[1]
    _, _, Vt = np.linalg.svd(A)
    X = Vt[:, 3, :]
    return X[:, :3] / X[:, 3:4]


class Triangulator(object):
    def __init__(self, p1, p2, f1, f2, c1, c2, dist1, dist2):
        self.f1, self.f2, self.c1, self.c2 = f1, f2, c1, c2
        self.dist1, self.dist2 = dist1, dist2
        self.R = None
        self.t = None
        self.p1 = self._normalize(p1, f1, c1, dist1)
        self.p2 = self._normalize(p2, f2, c2, dist2)

    @staticmethod
    def _normalize(p, f, c, dist):
        src = np.zeros((1, p.shape[0], 2), dtype=np.float32)
        src[0] = p
        K = np.asarray([[f, 0.0, c[0]], [0.0, f, c[1]], [0.0, 0.0, 1.0]])
        return cv2.undistortPoints(src, K, np.asarray(dist, dtype=np.float64), P=None).reshape(-1, 2)

    def getFundamentalMat(self):
        return cv2.findFundamentalMat(self.p2, self.p1, method=cv2.FM_8POINT)[0]

    def getQuaternion(self):
        if self.R is not None and self.R.any():
            return quaternions.quatFromRotationMatrix(gs(self.R)).asVector()

    def triangulate(self):
        F = self.getFundamentalMat()
        U, _, Vt = np.linalg.svd(F)
        W = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
        t = U[:, 2]
        cands = [U @ W.T @ Vt, U @ W @ Vt, U @ W.T @ (-Vt), U @ W @ (-Vt)]
        Rs2 = cands[:2]
        cnt = 0
        for R in cands:                         # keep the proper rotations (det != -1), as the reference does
            if np.round(np.linalg.det(R)) != -1:
                Rs2[cnt] = R
                cnt += 1
        P1 = np.hstack([np.eye(3), np.zeros((3, 1))])
        x1 = self.p1.astype(np.float64); x2 = self.p2.astype(np.float64)
        outs, plusses = [], []
        for R in Rs2:
            P = np.hstack([R, t.reshape(3, 1)])
            Pi = np.hstack([np.linalg.inv(R), (-t @ R).reshape(3, 1)])
            out = _triangulate_pairs(P, P1, x1, x2)
            out_ = _triangulate_pairs(P1, Pi, x1, x2)
            outs += [out, out_]
            plusses += [int(np.sum(np.where(out[:, 2] > 0, 1, -1))), int(np.sum(np.where(out_[:, 2] > 0, 1, -1)))]
        if abs(plusses[0]) > abs(plusses[2]):
            zdx = 0
        elif abs(plusses[0]) < abs(plusses[2]):
            zdx = 2
        else:
            zdx = 0 if abs(plusses[0] + plusses[1]) > abs(plusses[2] + plusses[3]) else 2
        self.R = Rs2[zdx // 2]
        self.t = t if plusses[zdx] < 0 else -t
        return outs[zdx] * -1 * np.sign(plusses[zdx])


class multiTriangulator(object):
    def __init__(self, pts, intrins):
        if type(pts) != np.ndarray:
            raise ArgusError("points passed must be a numpy array")
        if len(pts.shape) != 2:
            raise ArgusError("points must be a 2d array")
        self.pts = pts
        self.ncams = int(pts.shape[1] / 2)
        self.intrins = intrins
        self.ext = None
        if self.ncams != len(self.intrins):
            raise ArgusError("length of intrinsics does not match the number of cameras supplied")

    def triangulate(self):
        origin = self.pts[:, -2:]
        n = self.pts.shape[0]
        xyzs, transrots = [], []
        for k in range(self.ncams - 1):
            cam = self.pts[:, 2 * k:2 * k + 2]
            good = np.nonzero(~np.isnan(cam).any(axis=1) & ~np.isnan(origin).any(axis=1))[0]
            tring = Triangulator(cam[good], origin[good], self.intrins[k, 0], self.intrins[-1, 0], self.intrins[k, 1:3],
                                 self.intrins[-1, 1:3], self.intrins[k, -5:], self.intrins[-1, -5:])
            tri = tring.triangulate()
            transrots.append(np.hstack((tring.getQuaternion(), tring.t)))
            xyz = np.zeros((n, 3))
            xyz[good] = tri
            xyz[xyz == 0] = np.nan
            xyzs.append(xyz)
        self.ext = np.vstack(transrots + [[1, 0, 0, 0, 0, 0, 0]])
        return self.normalizeAndAverage(xyzs)

    def normalizeAndAverage(self, xyzs):
        import warnings
        dists = []
        for xyz in xyzs:
            ok = ~np.isnan(xyz).any(axis=1)
            dists.append(np.mean(np.linalg.norm(xyz[ok], axis=1)) if ok.any() else np.nan)
        ret = np.stack([xyzs[0]] + [xyzs[k] * (dists[0] / dists[k]) for k in range(1, len(xyzs))])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            return np.nanmean(ret, axis=0)
