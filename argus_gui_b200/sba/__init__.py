"""``sba``-compatible Python module backed by the B200 bundle-adjustment kernels.

Argus drives its bundle adjustment through the third-party package python-sba 1.6.8.2
(``sba @ git+https://github.com/backyardbiomech/python-sba.git@python310``, pinned at commit 9a431c51 in
/root/reference/uv.lock:2139-2142), which is a ctypes binding of Lourakis' sba 1.6 C library and is NOT part of the
reference tree.  This module re-provides exactly the names Argus uses from it
(/root/reference/argus_gui/sbaDriver.py:270-300, /root/reference/argus_gui/triangulate.py:7,65):

    sba.Points.fromDylan(ndarray)            rows [X Y Z u1 v1 ... um vm], NaN = not visible
    sba.Cameras.fromDylan(ndarray m x 17)    [fx cx cy AR s r2 r4 t1 t2 r6 qw qx qy qz tx ty tz]
    cameras.toTxt(path), points.toTxt(path)  np.loadtxt-readable
    sba.Options.fromInput(cameras, points)   with .camera, .nccalib, .ncdist (and ncon, mcon, maxIter, opts, verbose)
    sba.OPTS_CAMS, sba.OPTS_FIX{5,4,2,...}_INTR, sba.OPTS_FIX{5,4,3,...,0}_DIST
    sba.SparseBundleAdjust(cameras, points, options) -> (newcameras, newpoints, info)
    info.printResults()
    sba.quaternions.quatFromRotationMatrix(R).asVector()

so that ``import argus_gui_b200.sba as sba`` (or putting this package on the path as ``sba``) makes
``sbaArgusDriver.fix()`` run on the GPU unchanged.  Packing / unpacking follows the reference C driver that python-sba
mirrors: sba-1.6p/demo/readparams.c:161-224, 294-390 and eucsbademo.c:1503-1511 (local rotation starts at zero, the
initial rotation is a separate constant quaternion), 1652-1676 (recombination, sign so that qw >= 0).
Defaults that python-sba's unseen source decides are parameters here, with the C demo's values: 150 iterations,
opts = [1e-3, 1e-12, 1e-12, 1e-12, 0], ncon = mcon = 0, expert drivers, analytic Jacobian, no covariances.
"""
import sys
import numpy as np

from . import quaternions
from .. import ba as _ba

__all__ = ["Points", "Cameras", "Options", "Information", "SparseBundleAdjust", "quaternions",
           "OPTS_CAMS", "OPTS_CAMS_NODIST", "OPTS_NO_CAMS"]
__version__ = "1.6.8.2+b200"

# camera model selectors (Argus only ever sets OPTS_CAMS: intrinsics + distortion + pose, 16 + 1 values per camera)
OPTS_NO_CAMS = 0
OPTS_CAMS_NODIST = 1
OPTS_CAMS = 2
# number of trailing intrinsics kept fixed == struct globs_.nccalib (sba-1.6p/libsbaprojs/sbaprojs.h:131-140)
OPTS_FIX0_INTR = 0
OPTS_FIX1_INTR = 1
OPTS_FIX2_INTR = 2
OPTS_FIX3_INTR = 3
OPTS_FIX4_INTR = 4
OPTS_FIX5_INTR = 5
# number of trailing distortion coefficients kept fixed == struct globs_.ncdist (sbaprojs.h:142-151)
OPTS_FIX0_DIST = 0
OPTS_FIX1_DIST = 1
OPTS_FIX2_DIST = 2
OPTS_FIX3_DIST = 3
OPTS_FIX4_DIST = 4
OPTS_FIX5_DIST = 5

OPTS_MOTSTRUCT = 0
OPTS_MOT = 1
OPTS_STRUCT = 2

_FMT = "%.17g"   # enough digits for an exact double round trip through the text files Argus re-reads (sbaDriver.py:299-305)


class Points(object):
    """3-D points and their image projections.  ``B`` (n x 3), ``vmask`` (n x m uint8), ``X`` (nobs x 2, point-major,
    cameras ascending - the order of sba's measurement vector, sba_levmar.c:466-470)."""

    def __init__(self, B, vmask, X):
        self.B = np.ascontiguousarray(B, dtype=np.float64)
        self.vmask = np.ascontiguousarray(vmask, dtype=np.uint8)
        self.X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, 2)
        self.n, self.ncameras = self.vmask.shape
        if self.B.shape != (self.n, 3) or self.X.shape[0] != int(self.vmask.sum()):
            raise ValueError("inconsistent points / visibility / projections")

    @classmethod
    def fromDylan(cls, arr):
        arr = np.asarray(arr, dtype=np.float64)
        if arr.ndim != 2 or (arr.shape[1] - 3) % 2 != 0 or arr.shape[1] < 5:
            raise ValueError("expected rows [X Y Z u1 v1 ... um vm]")
        m = (arr.shape[1] - 3) // 2
        uv = arr[:, 3:].reshape(-1, m, 2)
        vis = ~np.isnan(uv).any(axis=2)
        return cls(arr[:, :3].copy(), vis.astype(np.uint8), uv[vis])

    @classmethod
    def fromTxt(cls, path, ncameras):
        """sba text format (sba-1.6p/demo/README.txt:9-30): X Y Z nframes frame0 x0 y0 frame1 x1 y1 ..."""
        B, rows = [], []
        with open(path) as f:
            for line in f:
                line = line.strip()
                if not line or line.startswith("#"):
                    continue
                t = [float(v) for v in line.split()]
                B.append(t[:3])
                nf = int(t[3])
                rows.append(sorted((int(t[4 + 3 * q]), t[5 + 3 * q], t[6 + 3 * q]) for q in range(nf)))
        vmask = np.zeros((len(B), ncameras), dtype=np.uint8); X = []
        for i, obs in enumerate(rows):
            for (j, u, v) in obs:
                vmask[i, j] = 1; X.append((u, v))
        return cls(np.array(B), vmask, np.array(X).reshape(-1, 2))

    def toDylan(self):
        out = np.full((self.n, 3 + 2 * self.ncameras), np.nan)
        out[:, :3] = self.B
        uv = out[:, 3:].reshape(self.n, self.ncameras, 2)
        uv[self.vmask != 0] = self.X
        return out

    def toTxt(self, path):
        np.savetxt(path, self.B, fmt=_FMT)

    def __len__(self):
        return self.n


class Cameras(object):
    """Camera rows [fx cx cy AR s | r2 r4 t1 t2 r6 | qw qx qy qz | tx ty tz] (m x 17)."""

    def __init__(self, arr):
        arr = np.array(arr, dtype=np.float64, copy=True)
        if arr.ndim != 2 or arr.shape[1] != 17:
            raise ValueError("expected an m x 17 camera array")
        self.arr = arr
        self.ncameras = arr.shape[0]

    @classmethod
    def fromDylan(cls, arr):
        return cls(arr)

    @classmethod
    def fromTxt(cls, path):
        """sba text formats (sba-1.6p/demo/README.txt:32-82): 17 values per line, or 12 (no distortion)."""
        rows = []
        with open(path) as f:
            for line in f:
                line = line.strip()
                if line and not line.startswith("#"):
                    rows.append([float(v) for v in line.split()])
        a = np.array(rows)
        if a.shape[1] == 12:
            a = np.hstack([a[:, :5], np.zeros((a.shape[0], 5)), a[:, 5:]])
        return cls(a)

    def toTxt(self, path):
        np.savetxt(path, self.arr, fmt=_FMT)

    def pack(self):
        """-> (a (m x 16 with the local rotation zeroed), rot0 (m x 4 unit, scalar >= 0)): readparams.c:161-224,
        eucsbademo.c:1503-1511."""
        m = self.ncameras
        q = self.arr[:, 10:14].copy()
        nrm = np.linalg.norm(q, axis=1, keepdims=True)
        q = q / nrm
        q[q[:, 0] < 0] *= -1
        a = np.zeros((m, 16))
        a[:, :10] = self.arr[:, :10]
        a[:, 13:16] = self.arr[:, 14:17]
        return a, q

    @staticmethod
    def unpack(a, rot0):
        """Recombine the refined local rotation with the initial one: q = (sqrt(1 - |v|^2), v) (x) q0, sign so that
        qw >= 0 (eucsbademo.c:1652-1676)."""
        m = a.shape[0]
        out = np.zeros((m, 17))
        out[:, :10] = a[:, :10]
        out[:, 14:17] = a[:, 13:16]
        v = a[:, 10:13]
        w = np.sqrt(np.maximum(0.0, 1.0 - (v * v).sum(axis=1)))
        b0, b1, b2, b3 = rot0[:, 0], rot0[:, 1], rot0[:, 2], rot0[:, 3]
        q = np.stack([w * b0 - v[:, 0] * b1 - v[:, 1] * b2 - v[:, 2] * b3,
                      w * b1 + v[:, 0] * b0 + v[:, 1] * b3 - v[:, 2] * b2,
                      w * b2 - v[:, 0] * b3 + v[:, 1] * b0 + v[:, 2] * b1,
                      w * b3 + v[:, 0] * b2 - v[:, 1] * b1 + v[:, 2] * b0], axis=1)
        q[q[:, 0] < 0] *= -1
        out[:, 10:14] = q
        return out

    def __len__(self):
        return self.ncameras


class Options(object):
    """Solver options.  ``nccalib`` / ``ncdist``: number of trailing intrinsics / distortion coefficients kept fixed."""

    def __init__(self, ncameras, npoints):
        self.ncameras, self.npoints = ncameras, npoints
        self.camera = OPTS_CAMS
        self.motstruct = OPTS_MOTSTRUCT
        self.nccalib = OPTS_FIX5_INTR
        self.ncdist = OPTS_FIX5_DIST
        self.ncon = 0                 # leading points kept fixed
        self.mcon = 0                 # leading cameras kept fixed
        self.maxIter = _ba.DEFAULT_ITMAX
        self.verbose = 0
        self.opts = list(_ba.DEFAULT_OPTS)     # [tau, eps1, eps2, eps3, eps4]
        self.expert = True
        self.analyticjac = True
        self.compat_sba16 = True      # reproduce sba 1.6's damping behaviour after a rejected step
        # EXTENSION, parity unpinned (python-sba / sba 1.6 have no such option): wand-length constraint inside the solver,
        # residual wand_weight * (wand_length - ||X_a - X_b||) per pair; wand_pairs = (ia, ib) point indices, off by default
        self.wand_pairs = None
        self.wand_length = None
        self.wand_weight = 0.0

    @classmethod
    def fromInput(cls, cameras, points):
        return cls(len(cameras), len(points))


class Information(object):
    """The info[10] vector of sba_motstr_levmar_x (sba_levmar.c:503-518) with names."""
    _REASONS = {1: "stopped by small ||J^T e||", 2: "stopped by small ||delta||", 3: "stopped by maxiter",
                4: "stopped by small relative reduction in ||e||", 5: "stopped by small ||e||",
                6: "stopped due to excessive failed attempts to increase damping", 7: "stopped due to infinite values in the functions"}

    def __init__(self, info, retval, nobs, nvars):
        self.info = np.asarray(info, dtype=np.float64)
        self.retval = int(retval)
        self.nprojs = int(nobs)
        self.nvars = int(nvars)
        self.initialError, self.finalError = self.info[0], self.info[1]
        self.iterations = int(self.info[5]); self.reason = int(self.info[6])
        self.funcEvals, self.jacEvals, self.linSystems = int(self.info[7]), int(self.info[8]), int(self.info[9])

    def printResults(self):
        n = max(self.nprojs, 1)
        print("SBA returned %d in %d iter, reason %d (%s), error %g [initial %g], %d/%d func/fjac evals, %d lin. systems" % (
            self.retval, self.iterations, self.reason, self._REASONS.get(self.reason, "?"), self.finalError / n, self.initialError / n,
            self.funcEvals, self.jacEvals, self.linSystems))
        sys.stdout.flush()


def SparseBundleAdjust(cameras, points, options):
    """Bundle adjustment of ``cameras`` / ``points``; returns (newcameras, newpoints, info).  Inputs are not modified.
    ``options.motstruct`` selects the driver as in the reference's demo (sba-1.6p/demo/eucsbademo.c:1560-1602):
    OPTS_MOTSTRUCT (sba_motstr_levmar_x + KDRTS callbacks, what Argus uses), OPTS_MOT (cameras only, sba_mot_levmar_x + KDRT)
    or OPTS_STRUCT (points only, sba_str_levmar_x + KDS)."""
    if options.camera != OPTS_CAMS:
        raise NotImplementedError("only the full Argus camera model (OPTS_CAMS: intrinsics + distortion + pose) is supported")
    if options.motstruct not in (OPTS_MOTSTRUCT, OPTS_MOT, OPTS_STRUCT):
        raise ValueError("options.motstruct must be OPTS_MOTSTRUCT, OPTS_MOT or OPTS_STRUCT")
    if points.ncameras != cameras.ncameras:
        raise ValueError("points were built for %d cameras, got %d" % (points.ncameras, cameras.ncameras))
    a, rot0 = cameras.pack()
    m, n = cameras.ncameras, points.n
    common = dict(itmax=options.maxIter, opts=options.opts, verbose=options.verbose, compat=options.compat_sba16)
    if options.motstruct == OPTS_MOT:
        cams, info, ret = _ba.solve_mot(points.vmask, a.ravel(), points.B.ravel(), rot0.ravel(), points.X, options.nccalib, options.ncdist,
                                        mcon=options.mcon, **common)
        pn = np.concatenate([cams, points.B.ravel()])
        nvars = 16 * m
    elif options.motstruct == OPTS_STRUCT:
        pts, info, ret = _ba.solve_str(points.vmask, points.B.ravel(), a.ravel(), rot0.ravel(), points.X, ncon=options.ncon, **common)
        pn = np.concatenate([a.ravel(), pts])
        nvars = 3 * n
    elif options.wand_pairs is not None and options.wand_weight > 0 and len(options.wand_pairs[0]) > 0:
        # constrained problem: the resident-data calls (the one-shot call is the reference's signature and has no pair list)
        p = np.concatenate([a.ravel(), points.B.ravel()])
        B = _ba.BundleAdjuster().upload(points.vmask, p, rot0.ravel(), points.X, options.nccalib, options.ncdist, ncon=options.ncon, mcon=options.mcon)
        try:
            B.set_wand(options.wand_pairs[0], options.wand_pairs[1], options.wand_length, options.wand_weight)
            ret, info = B.solve(itmax=options.maxIter, opts=options.opts, verbose=options.verbose, compat=options.compat_sba16)
            pn = B.download()
        finally:
            B.close()
        nvars = 16 * m + 3 * n
    else:
        p = np.concatenate([a.ravel(), points.B.ravel()])
        pn, info, ret = _ba.solve_motstr(points.vmask, p, rot0.ravel(), points.X, options.nccalib, options.ncdist, ncon=options.ncon,
                                         mcon=options.mcon, **common)
        nvars = 16 * m + 3 * n
    newcams = Cameras(Cameras.unpack(pn[:16 * m].reshape(m, 16), rot0))
    newpts = Points(pn[16 * m:].reshape(n, 3), points.vmask, points.X)
    return newcams, newpts, Information(info, ret, points.X.shape[0], nvars)
