"""Headless drop-in for the wand-calibration driver /root/reference/argus_gui/sbaDriver.py.

``sbaArgusDriver`` keeps the reference's constructor signature, attributes and data conventions
(sbaDriver.py:50-77): the best camera is swapped to the LAST column (``rearrange``, :118-162, the
caller's arrays are modified in place exactly as the reference does), rows are kept only if the reference camera and at
least one other camera see them (``parse``, :81-100), points are stacked [reference ; paired end 1 ; paired end 2 ;
unpaired tracks] (``getPointsAndExtArray``, :165-257), the initial structure / extrinsics come from
``multiTriangulator``, and ``fix()`` (:264-415) maps ``modeString`` to nccalib / ncdist, runs
``sba.SparseBundleAdjust`` - here the CUDA solver - and writes the same ``-sba-profile.txt``, ``-clicker-profile.txt``,
``-camera-N-profile.txt``, ``-dlt-coefficients.csv`` and ``-(un)paired-points-xyz.csv`` files.

The Qt / OpenGL viewer (``OutlierWindow`` is a QWidget in the reference) is out of scope; its numerical part
(``buildData`` :676-789, ``getCoefficients`` :1027-1165, ``pairedIsomorphism`` :988-1009, ``WandOutputter``
output.py:29-66) is provided by ``WandResults`` without any GUI toolkit, including the alignment of the scaled cloud to
user reference points (``transform`` :828-985: 1- to 4-point axis references, gravity, horizontal plane) and the sign change of
the z-related DLT coefficients the reference applies for axis references (:767-773).
"""
from __future__ import print_function

import copy
import random
import string
import sys

import numpy as np
import pandas

from . import sba
from . import dlt as _dlt
from .tools import undistort_pts, redistort_pts
from .triangulate import multiTriangulator

__all__ = ["sbaArgusDriver", "WandResults", "align_to_reference"]


def _rotation_between(a, b):
    """Rotation that turns direction ``a`` into direction ``b`` (Rodrigues' formula in its cross-product-matrix form; the
    reference's ``rotation`` helper, sbaDriver.py:1232-1248)."""
    a = a / np.linalg.norm(a)
    b = b / np.linalg.norm(b)
    v = np.cross(a, b)
    K = np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])
    return np.eye(3) + K + (K @ K) * ((1.0 - np.dot(a, b)) / np.dot(v, v))


def _kabsch(A, B):
    """Least-squares rotation taking the centred point set A onto B (the reference's ``rigid_transform_3D``,
    sbaDriver.py:1271-1299, which flips the last right singular vector when the solution is a reflection)."""
    Ac = A - A.mean(axis=0)
    Bc = B - B.mean(axis=0)
    U, _, Vt = np.linalg.svd(Ac.T @ Bc)
    R = Vt.T @ U.T
    if np.linalg.det(R) < 0:
        print('Reflection detected - correcting to ensure proper right-handed coordinate system')
        Vt[2, :] *= -1
        R = Vt.T @ U.T
    return R


def align_to_reference(xyzs, nref, reference_type='Axis points', recording_frequency=100):
    """The reference's ``OutlierWindow.transform`` (sbaDriver.py:828-985): move / rotate the scaled cloud ``xyzs`` (whose first
    ``nref`` rows are the reference points) into the user's frame.

    * 'Axis points': 1 point = origin; 2 = origin, +z; 3 = origin, +x, +y (z = -(x cross y), as the reference has it);
      4 = origin, +x, +y, +z.  Any other count only moves the origin.
    * 'Gravity': the reference rows are a free-fall track; a quadratic per coordinate gives the acceleration, which is turned
      onto -z.
    * 'Plane': the reference rows lie in a horizontal plane; principal axes from an SVD, +z towards the bulk of the data, origin
      at the centre of the reference points.
    """
    xyzs = np.asarray(xyzs, dtype=np.float64)
    ref = xyzs[:nref]
    origin = ref[0].copy()
    out = xyzs - origin
    ref = ref - origin
    k = ref.shape[0]
    if reference_type == 'Axis points' and k == 1:
        print('Using 1-point (origin) reference axes')
    elif reference_type == 'Axis points' and k == 2:
        print('Using 2-point (origin,+Z) reference axes')
        out = out @ _rotation_between(ref[1], np.array([0.0, 0.0, 1.0])).T
    elif reference_type == 'Axis points' and k in (3, 4):
        if k == 3:
            print('Using 3-point (origin,+x,+y) reference axes')
            z = -np.cross(ref[1], ref[2])
            A = np.vstack((ref, z / np.linalg.norm(z)))
            lz = 1.0
        else:
            print('Using 4-point (origin,+x,+y,+z) reference axes')
            A = ref
            lz = np.linalg.norm(A[3])
        B = np.diag([np.linalg.norm(A[1]), np.linalg.norm(A[2]), lz])
        out = out @ _kabsch(A, np.vstack((np.zeros(3), B))).T
    elif reference_type == 'Gravity':
        print('Using gravity alignment, +Z will point anti-parallel to gravity')
        f = float(recording_frequency)
        t = np.arange(k)
        ok = np.isfinite(ref[:, 0])
        acc = np.array([np.mean(np.diff(np.polyval(np.polyfit(t[ok], ref[ok, a], 2), t), n=2)) for a in range(3)]) * f * f
        g = acc / np.linalg.norm(acc)
        down = np.array([0.0, 0.0, -1.0])
        axis = np.cross(g, down)
        axis = axis / np.linalg.norm(axis)
        ang = np.arccos(np.dot(g, down))
        K = np.array([[0.0, -axis[2], axis[1]], [axis[2], 0.0, -axis[0]], [-axis[1], axis[0], 0.0]])
        R = np.eye(3) + np.sin(ang) * K + (1.0 - np.cos(ang)) * (K @ K)
        out = out @ R.T
        print('Gravity measured with {:.2f}% accuracy!'.format(np.linalg.norm(acc) / 9.81 * 100))
    elif reference_type == 'Plane':
        print('Aligning to horizontal plane reference points')
        avg = ref.mean(axis=0)
        basis = np.linalg.svd(ref - avg, full_matrices=True)[2].T
        if np.linalg.det(basis) == -1:
            basis = -basis
        # the reference centres the UNSHIFTED cloud on the mean of the shifted reference points here (sbaDriver.py:953); the
        # final re-centring on the reference points makes the difference a pure translation that cancels
        centred = xyzs - avg
        out = centred @ basis
        if out[:, 2].mean() < 0:
            basis = basis @ np.diag([1.0, -1.0, -1.0])
            out = centred @ basis
        out = out - out[:k].mean(axis=0)
    return out


class sbaArgusDriver(object):
    def __init__(self, ppts, uppts, cams, display=True, scale=None, modeString=None, ref=None, name=None, temp="",
                 report=True, outputCPs=True, reorder=True, reference_type='Axis points', recording_frequency=100, wand_weight=0.0):
        """The reference's constructor (sbaDriver.py:50-77) plus ONE optional extension at the end: ``wand_weight`` > 0 puts
        the wand length into the bundle adjustment itself (parity unpinned; residual = wand_weight pixels per unit RELATIVE
        length error of every paired frame); 0 is the reference's behaviour, where the wand only sets the scale afterwards."""
        self.ppts = ppts
        self.uppts = uppts
        self.cams = cams
        self.ncams = cams.shape[0]
        self.display = display
        self.scale = scale
        self.modeString = modeString
        self.nppts = 0
        self.nuppts = 0
        self.ref = ref
        self.name = name
        self.temp = temp
        self.report = report
        self.outputCameraProfiles = outputCPs
        self.reorder = reorder
        self.reference_type = reference_type
        self.recording_frequency = recording_frequency
        self.wand_weight = float(wand_weight)
        self.sba_options = {}           # extra solver options (maxIter, opts, mcon, ...) applied in fix()
        self.info = None
        self.grapher = None
        self.outliers = []
        print('Parsing points...')
        sys.stdout.flush()
        if self.reorder:
            self.rearrange()
        else:
            self.order = np.arange(self.ncams)
        self.pts, self.ext, self.indices = self.getPointsAndExtArray()

    # ---- sbaDriver.py:81-100 ----
    def parse(self, pts):
        if hasattr(pts, "toarray"):
            pts = pts.toarray()
        ret = []
        nc2 = 2 * self.ncams
        for k in range(int(pts.shape[1] / nc2)):
            p = pts[:, k * nc2:(k + 1) * nc2]
            nan = np.isnan(p)
            bad = nan[:, nc2 - 2:].any(axis=1) | nan[:, :nc2 - 2].all(axis=1)
            good = np.nonzero(~bad)[0]
            ret.append([p[good], good])
        return ret

    # ---- sbaDriver.py:104-115 ----
    def count(self, pts):
        if hasattr(pts, "toarray"):
            pts = pts.toarray()
        counts = np.zeros(self.ncams)
        nc2 = 2 * self.ncams
        for j in range(int(pts.shape[1] / nc2)):
            nan = np.isnan(pts[:, j * nc2:(j + 1) * nc2])
            others = ((~nan).sum(axis=1) / 2).astype(np.int64) - 1       # int(count(False) / 2) - 1
            for i in range(self.ncams):
                counts[i] += others[~nan[:, 2 * i]].sum()
        return counts

    # ---- sbaDriver.py:118-162 ----
    def rearrange(self):
        counts = np.zeros(self.ncams)
        if self.ppts is not None:
            counts += self.count(self.ppts)
        if self.uppts is not None:
            counts += self.count(self.uppts)
        best = int(np.argmax(counts))
        print('Using camera ' + str(best + 1) + ' as reference')
        sys.stdout.flush()
        nc = self.ncams
        self.order = np.arange(nc)
        if best != nc - 1:
            def swap(arr, a, b):
                tmp = copy.copy(arr[:, a:a + 2]); arr[:, a:a + 2] = arr[:, b:b + 2]; arr[:, b:b + 2] = tmp
            if self.ppts is not None:
                swap(self.ppts, 2 * (nc - 1), 2 * best)
                swap(self.ppts, 2 * nc + 2 * (nc - 1), 2 * nc + 2 * best)
            if self.uppts is not None:
                for k in range(int(self.uppts.shape[1] / (2 * nc))):
                    swap(self.uppts, 2 * nc * k + 2 * (nc - 1), 2 * nc * k + 2 * best)
            if self.ref is not None:
                swap(self.ref, 2 * (nc - 1), 2 * best)
            self.order[best] = nc - 1
            self.order[-1] = best
            tmp = copy.copy(self.cams[-1, :]); self.cams[-1, :] = self.cams[best, :]; self.cams[best, :] = tmp

    # ---- sbaDriver.py:165-257 ----
    def getPointsAndExtArray(self):
        indices = {'paired': None, 'unpaired': None}
        stacks = []
        if self.ppts is not None:
            parsed = self.parse(self.ppts)
            paired = np.vstack([q[0] for q in parsed])
            indices['paired'] = [list(q[1]) for q in parsed]
            self.nppts = paired.shape[0]
            stacks.append(paired)
        if self.uppts is not None:
            parsed = self.parse(self.uppts)
            unpaired = np.vstack([q[0] for q in parsed])
            indices['unpaired'] = [list(q[1]) for q in parsed]
            self.nuppts = unpaired.shape[0]
            stacks.append(unpaired)
        pts = np.vstack(stacks)
        if self.ref is not None:
            print('Including reference points')
            sys.stdout.flush()
            pts = np.vstack((self.ref, pts))
        print('Triangulating...')
        sys.stdout.flush()
        tring = multiTriangulator(pts, self.cams)
        xyz = tring.triangulate()
        return np.hstack((xyz, pts)), tring.ext, indices

    def id_generator(self, size=12, chars=string.ascii_uppercase + string.digits):
        return ''.join(random.choice(chars) for _ in range(size))

    # ---- sbaDriver.py:264-415 ----
    def fix(self):
        print('Found ' + str(self.pts.shape[0]) + ' 3D points')
        print('Passing points to SBA...')
        sys.stdout.flush()
        points = sba.Points.fromDylan(self.pts)
        cameras = sba.Cameras.fromDylan(np.hstack((self.cams, self.ext)))
        cameras.toTxt(self.name + '-sba-profile-orig.txt')
        options = sba.Options.fromInput(cameras, points)
        options.camera = sba.OPTS_CAMS
        options.nccalib = {'0': sba.OPTS_FIX5_INTR, '1': sba.OPTS_FIX4_INTR, '2': sba.OPTS_FIX2_INTR}.get(self.modeString[0], options.nccalib)
        options.ncdist = {'0': sba.OPTS_FIX5_DIST, '1': sba.OPTS_FIX4_DIST, '2': sba.OPTS_FIX3_DIST,
                          '3': sba.OPTS_FIX0_DIST}.get(self.modeString[1], options.ncdist)
        if self.wand_weight > 0 and self.indices.get('paired') is not None:
            # the frames seen in both tracks are the wand pairs (pairedIsomorphism); the solver's length unit is the initial
            # triangulation's, so the known length is expressed as the current mean and the true scale is applied afterwards as ever
            i0, i1 = self.indices['paired']
            pos0 = {f: k for k, f in enumerate(i0)}; pos1 = {f: k for k, f in enumerate(i1)}
            common = sorted(set(i0) & set(i1))
            nref = self.ref.shape[0] if self.ref is not None else 0
            ia = np.array([nref + pos0[f] for f in common], dtype=np.int64)
            ib = np.array([nref + len(i0) + pos1[f] for f in common], dtype=np.int64)
            if ia.size:
                L0 = float(np.linalg.norm(points.B[ia] - points.B[ib], axis=1).mean())
                options.wand_pairs, options.wand_length, options.wand_weight = (ia, ib), L0, self.wand_weight / L0
        for k, v in self.sba_options.items():
            setattr(options, k, v)
        newcameras, newpoints, info = sba.SparseBundleAdjust(cameras, points, options)
        info.printResults()
        sys.stdout.flush()
        self.info = info
        self.newcameras, self.newpoints = newcameras, newpoints

        key = self.id_generator()
        if self.temp:
            newcameras.toTxt(self.temp + '/' + key + '_cn.txt')
            newpoints.toTxt(self.temp + '/' + key + '_np.txt')
        camO = newcameras.arr.copy()
        if self.order is not None:
            camO = camO[list(self.order), :]
        pandas.DataFrame(camO).to_csv(self.name + '-sba-profile.txt', header=False, index=False, sep=' ')
        clicker = [[i + 1, camO[i, 0], 2 * camO[i, 1], 2 * camO[i, 2], camO[i, 1], camO[i, 2], camO[i, 3], camO[i, 5], camO[i, 6],
                    camO[i, 7], camO[i, 8], camO[i, 9]] for i in range(len(camO))]
        pandas.DataFrame(clicker).to_csv(self.name + '-clicker-profile.txt', header=False, index=False, sep=' ')

        npframes = self.ppts.shape[0] if self.ppts is not None else None
        nupframes = self.uppts.shape[0] if self.uppts is not None else None
        uvs = [undistort_pts(np.ascontiguousarray(self.pts[:, 3 + 2 * k:3 + 2 * (k + 1)]), self.cams[k]) for k in range(self.ncams)]
        nRef = self.ref.shape[0] if self.ref is not None else 0
        self.grapher = WandResults(newpoints.B, newcameras.arr, self.nppts, self.nuppts, self.scale, self.ref is not None, self.indices,
                                   self.ncams, npframes, nupframes, self.name, uvs, nRef, self.order, self.cams,
                                   self.reference_type, self.recording_frequency)
        self.outliers = self.grapher.outliers
        if self.outputCameraProfiles:
            nps = np.loadtxt(self.name + '-sba-profile.txt', ndmin=2)
            for k in range(len(nps)):
                l = np.insert(nps[k], 0, 1.)
                l = np.delete(l, [5] + list(np.arange(11, 18)), axis=0)
                np.savetxt(self.name + '-camera-' + str(self.order[k] + 1) + '-profile.txt', np.asarray([l]), fmt='%-1.5g')
        if self.report:
            print('Found ' + str(len(self.outliers)) + ' possible outliers')
        return self.grapher

    def showGraph(self):
        """The reference opens a Qt/OpenGL window here; the headless driver has nothing to show."""
        return None

    def exitLoop(self):
        return None

    # ---- sbaDriver.py:425-481 ----
    def redo(self):
        self.pts = np.delete(self.pts, self.grapher.index, axis=0)
        tmp = self.indices['paired']
        if tmp is not None:
            for o in self.outliers:
                if o.get('kind') != 'paired':
                    continue
                lst = tmp[0] if o.get('paired_point') == 1 else tmp[1]
                if (o['frame'] - 1) in lst:
                    lst.remove(o['frame'] - 1)
                    self.nppts -= 1
        tmp = self.indices['unpaired']
        if tmp is not None:
            for o in self.outliers:
                if o.get('kind') == 'unpaired' and (o['frame'] - 1) in tmp[o['unpaired_track']]:
                    tmp[o['unpaired_track']].remove(o['frame'] - 1)
                    self.nuppts -= 1
        print('\nRunning again with outliers removed...')
        sys.stdout.flush()
        return self.fix()


class WandResults(object):
    """Numerical part of the reference's OutlierWindow (no GUI): wand scale and score, centring, per-camera 11-parameter
    DLT fit with reprojection RMSE and 3-sigma outliers, output files."""

    def __init__(self, xyzs, cam, nppts, nuppts, scale, refBool, indices, ncams, npframes, nupframes, name, uvs, nRef, order, cams,
                 reference_type='Axis points', recording_frequency=100):
        self.nppts, self.nuppts, self.scale, self.refBool = nppts, nuppts, scale, refBool
        self.indices, self.ncams, self.npframes, self.nupframes = indices, ncams, npframes, nupframes
        self.name, self.uvs, self.nRef, self.order = name, uvs, nRef, order
        self.reference_type, self.recording_frequency = reference_type, recording_frequency
        # 8-value pinhole profiles [f, cx, cy, k1, k2, p1, p2, k3] for the outlier report (sbaDriver.py:511-513)
        self.cams = [[c[u] for u in (0, 1, 2, 5, 6, 7, 8, 9)] for c in cams]
        self.cam = cam
        self.ref = None
        self.outliers, self.index = self.buildData(np.array(xyzs, dtype=np.float64, copy=True))

    def isVisible(self):
        return False

    def averDist(self, pts1, pts2):
        d = np.linalg.norm(pts1 - pts2, axis=1)
        return np.nanmean(d), np.nanstd(d)

    def pairedIsomorphism(self, pts):
        i0, i1 = self.indices['paired'][0], self.indices['paired'][1]
        n1 = len(i0)
        p1, p2 = pts[:n1, :], pts[n1:, :]
        common = sorted(set(i0) & set(i1))
        pos0 = {f: k for k, f in enumerate(i0)}; pos1 = {f: k for k, f in enumerate(i1)}
        set1 = np.asarray([p1[pos0[f]] for f in common]); set2 = np.asarray([p2[pos1[f]] for f in common])
        return p1, p2, set1, set2

    def buildData(self, xyzs):
        p1 = p2 = None
        factor = 1.0
        dist = std = None
        if self.nppts != 0:
            paired = xyzs[self.nRef:self.nppts + self.nRef]
            p1, p2, s1, s2 = self.pairedIsomorphism(paired)
            dist, std = self.averDist(s1, s2)
            factor = self.scale / dist
        xyzs = xyzs * factor
        if self.refBool:
            print('Using reference points')
            xyzs = align_to_reference(xyzs, self.nRef, self.reference_type, self.recording_frequency)
            self.ref = xyzs[:self.nRef, :]
        else:
            print('No reference points available - centering the calibration on the mean point location.')
            xyzs = xyzs - np.mean(xyzs, axis=0)
        if self.nppts != 0:
            paired = xyzs[self.nRef:self.nppts + self.nRef]
            p1, p2, self.pairedSet1, self.pairedSet2 = self.pairedIsomorphism(paired)
        self.up = xyzs[self.nRef + self.nppts:, :] if self.nuppts != 0 else None
        self.xyzs = xyzs
        dlts, errs, outliers, ptsi = [], [], [], []
        for camn, uv in enumerate(self.uvs):
            cos, error, outlier, ind = self.getCoefficients(xyzs, uv, camn)
            outliers += outlier
            ptsi += [i for i in ind if i not in ptsi]
            dlts.append(cos); errs.append(error)
        self.dlts = np.asarray(dlts)
        self.dlterrors = np.asarray(errs)
        if self.refBool and self.reference_type == 'Axis points':
            self.dlts[:, [2, 6, 10]] *= -1.0          # the z-related coefficients change sign with axis references (sbaDriver.py:767-773)
        self.outputDLT(self.dlts, self.dlterrors)
        if self.nppts != 0:
            self.wandscore = 100. * (std / dist)
            print('\nWand score: ' + str(self.wandscore))
        else:
            self.wandscore = 'not applicable'
            print('\nWand score: not applicable')
        sys.stdout.flush()
        self.writePoints(p1, p2)
        return outliers, ptsi

    def outputDLT(self, cos, error):
        print('\nDLT Errors: ')
        print(error)
        dat = np.asarray(cos).T[0]
        if self.order is not None:
            dat = dat[:, list(self.order)]
        pandas.DataFrame(dat).to_csv(self.name + '-dlt-coefficients.csv', header=False, index=False)

    def getCoefficients(self, xyz, uv, cam_index=0):
        uv = np.asarray(uv, dtype=np.float64)
        idx = np.nonzero(~np.isnan(uv[:, 0]))[0]
        n = len(idx)
        # least squares + reprojection distances on the GPU (argus_dlt_fit); the index bookkeeping below stays on the host
        Lv, err_all, (n_used, esum, merr, stderr) = _dlt.fit(np.asarray(xyz, dtype=np.float64), uv)
        L = Lv.reshape(11, 1)
        errors = err_all[idx]
        camera_number = (self.order[cam_index] + 1) if self.order is not None else cam_index + 1
        pb_1 = len(self.indices['paired'][0]) if self.indices['paired'] is not None else -1
        upidx = np.hstack(self.indices['unpaired']) if self.indices['unpaired'] is not None else None
        outliers, ptsi = [], []
        for k in np.nonzero(errors >= 3 * stderr + merr)[0]:
            i = int(idx[k])
            if i >= self.nRef and i not in ptsi:
                ptsi.append(i)
            base = {'uv': redistort_pts(np.array([uv[i]], dtype=np.float64), self.cams[cam_index])[0], 'error': float(errors[k]),
                    'camera': camera_number}
            if self.nRef - 1 < i < pb_1 + self.nRef:
                outliers.append(dict(base, frame=self.indices['paired'][0][i - self.nRef] + 1, kind='paired', paired_point=1, unpaired_track=None))
            elif pb_1 + self.nRef <= i < self.nppts + self.nRef:
                outliers.append(dict(base, frame=self.indices['paired'][1][i - pb_1 - self.nRef] + 1, kind='paired', paired_point=2,
                                     unpaired_track=None))
            elif self.nppts + self.nRef <= i < self.nuppts + self.nppts + self.nRef and upidx is not None:
                j = i - self.nppts - self.nRef
                acc = 0
                for t, lst in enumerate(self.indices['unpaired']):
                    if j < acc + len(lst):
                        outliers.append(dict(base, frame=int(upidx[j]) + 1, kind='unpaired', paired_point=None, unpaired_track=t))
                        break
                    acc += len(lst)
        rmse = float(esum / float(n)) if n else float('nan')
        return L, rmse, outliers, ptsi

    def writePoints(self, p1, p2):
        """output.py:29-66 (WandOutputter.output), vectorised."""
        if self.npframes is not None and p1 is not None:
            pout = np.full((self.npframes, 6), np.nan)
            pout[self.indices['paired'][0], :3] = p1
            pout[self.indices['paired'][1], 3:] = p2
            pout[pout == 0] = np.nan
            pandas.DataFrame(pout, columns='x_1,y_1,z_1,x_2,y_2,z_2'.split(',')).to_csv(self.name + '-paired-points-xyz.csv', index=False, na_rep='NaN')
        if self.up is not None and self.indices['unpaired'] is not None:
            nt = len(self.indices['unpaired'])
            upout = np.full((self.nupframes, 3 * nt), np.nan)
            start = 0
            for k, lst in enumerate(self.indices['unpaired']):
                upout[lst, 3 * k:3 * k + 3] = self.up[start:start + len(lst)]
                start += len(lst)
            upout[upout == 0] = np.nan
            cols = []
            for k in range(nt):
                cols += 'x_{0},y_{0},z_{0}'.format(k + 1).split(',')
            pandas.DataFrame(upout, columns=cols).to_csv(self.name + '-unpaired-points-xyz.csv', index=False, na_rep='NaN')
