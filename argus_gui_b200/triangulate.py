"""Initial structure / pose estimate that feeds the bundle adjustment: same interface as
/root/reference/argus_gui/triangulate.py (``multiTriangulator(pts, intrins).triangulate()`` -> xyz, attribute ``.ext``
= m x 7 [qw qx qy qz tx ty tz], last camera = reference frame).  SURVEY.md section 8 row f1: the per-point work of every
camera pair runs on the GPU (C ABI ``argus_pair_triangulate``, csrc/pair_kernels.cu); the cross-pair rescale / average is numpy.

Per camera k paired with the last (reference) camera (triangulate.py:27-212): undistort + normalise both pixel sets
(cv2.undistortPoints, float32), 8-point fundamental matrix, SVD -> two candidate rotations x two baselines, two-view
triangulation of every point for the forward and the inverse view, cheirality vote, sign of the baseline.  Across pairs the
point clouds are rescaled to the first pair's mean range and nan-averaged (triangulate.py:292-307).
"""
import ctypes
import numpy as np

from . import _lib
from .sba import quaternions
from .tools import ArgusError

__all__ = ["gs", "Triangulator", "multiTriangulator"]


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


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


class Triangulator(object):
    """One camera paired with the reference camera (triangulate.py:27-212).  ``triangulate()`` runs the whole pair on the GPU
    (``argus_pair_triangulate``: undistortion, 8-point fundamental matrix, candidate poses, two-view triangulation of every
    point under the four hypotheses, cheirality vote) and leaves ``R``, ``t`` like the reference."""

    def __init__(self, p1, p2, f1, f2, c1, c2, dist1, dist2):
        self.f1, self.f2, self.c1, self.c2 = f1, f2, c1, c2
        self.dist1, self.dist2 = dist1, dist2
        self.R = None
        self.t = None
        self.F = None
        self.p1 = np.ascontiguousarray(p1, dtype=np.float64).reshape(-1, 2)
        self.p2 = np.ascontiguousarray(p2, dtype=np.float64).reshape(-1, 2)

    def _profile(self, f, c, dist):
        return np.ascontiguousarray(np.concatenate([[f, c[0], c[1]], np.asarray(dist, dtype=np.float64)[-5:]]), dtype=np.float64)

    def getQuaternion(self):
        if self.R is not None and self.R.any():
            return quaternions.quatFromRotationMatrix(gs(self.R)).asVector()

    def triangulate(self):
        lib = _lib.load()
        n = self.p1.shape[0]
        if n < 8 or self.p2.shape[0] != n:
            raise ArgusError("at least 8 point pairs seen by both cameras are needed")
        xyz = np.empty((n, 3)); R = np.empty((3, 3)); t = np.empty(3); F = np.empty((3, 3))
        pr1 = self._profile(self.f1, self.c1, self.dist1); pr2 = self._profile(self.f2, self.c2, self.dist2)
        rc = lib.argus_pair_triangulate(_ptr(self.p1), _ptr(self.p2), n, _ptr(pr1), _ptr(pr2), _ptr(xyz), _ptr(R), _ptr(t), _ptr(F))
        _lib.check(rc, "argus_pair_triangulate")
        self.R, self.t, self.F = R, t, F
        return xyz

    def getFundamentalMat(self):
        if self.F is None:
            self.triangulate()
        return self.F


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
