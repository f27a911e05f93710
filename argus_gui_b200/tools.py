"""Drop-in for the DLT functions of /root/reference/argus_gui/tools.py (same names, arguments, return shapes, errors);
the per-frame Python loops of the reference are replaced by the batched CUDA kernels behind the C ABI.

    uv_to_xyz(pts, profs, dlt)            tools.py:296-337   -> argus_dlt_triangulate
    get_repo_errors(xyzs, pts, prof, dlt) tools.py:360-423   -> argus_dlt_reproj_error
    undistort_pts(pts, prof)              tools.py:129-148   -> argus_undistort_points   (pinhole profiles)
    solve_dlt(xyz, uv)                    tools.py:220-280   -> argus_dlt_fit
    reconstruct_uv(L, xyz), dlt_inverse(L, xyz), normalize(pts, prof)   (host numpy, not hot)

Only pinhole profiles (list / ndarray) are supported; the omnidirectional model objects of the external ``argus.ocam``
package are out of scope."""
import numpy as np

from . import dlt as _dlt

__all__ = ["ArgusError", "dlt_inverse", "reconstruct_uv", "undistort_pts", "normalize", "solve_dlt", "uv_to_xyz",
           "get_repo_errors", "uv_to_xyz_and_errors", "bootstrapXYZs"]


class Error(Exception):
    pass


class ArgusError(Error):
    """Exception raised for errors in the input (tools.py:17-29)."""

    def __init__(self, value):
        self.value = value

    def __str__(self):
        return repr(self.value)


def _is_pinhole(prof):
    return isinstance(prof, (list, tuple, np.ndarray))


def dlt_inverse(L, xyz):
    if not hasattr(L, "__iter__"):
        raise ArgusError("DLT coefficients must be iterable")
    L = np.asarray(L, dtype=np.float64)
    if L.ndim != 1 or L.shape[0] != 11:
        raise ArgusError("There must be exaclty 11 DLT coefficients in a 1d iterable")
    if type(xyz) != np.ndarray or xyz.ndim != 2 or xyz.shape[1] != 3:
        raise ArgusError("XYZ must be an Nx3 numpy array")
    den = xyz @ L[8:11] + 1.0
    return np.stack([(xyz @ L[0:3] + L[3]) / den, (xyz @ L[4:7] + L[7]) / den], axis=1)


def reconstruct_uv(L, xyz):
    if not hasattr(L, "__iter__"):
        raise ArgusError("DLT coefficients must be iterable")
    if type(xyz) != np.ndarray:
        raise ArgusError("XYZ must be a numpy array")
    L = np.asarray(L, dtype=np.float64)
    den = np.dot(L[-3:], xyz) + 1.0
    return np.array([(np.dot(L[:3], xyz) + L[3]) / den, (np.dot(L[4:7], xyz) + L[7]) / den])


def undistort_pts(pts, prof):
    """Nx2 pixel coordinates -> Nx2 undistorted pixel coordinates (float32, as cv2.undistortPoints returns them)."""
    if not _is_pinhole(prof):
        raise ArgusError("only pinhole distortion profiles are supported by the B200 backend")
    pts = np.asarray(pts, dtype=np.float64)
    if pts.ndim == 1:          # the reference broadcasts a (2,) pair into two identical rows (tools.py:138-139)
        pts = np.tile(pts.reshape(1, 2), (pts.shape[0], 1))
    return _dlt.undistort(pts, np.asarray(prof, dtype=np.float64))


def normalize(pts, prof):
    if type(pts) != np.ndarray:
        raise ArgusError("pts must be a numpy array")
    if pts.ndim != 2 or pts.shape[1] != 2:
        raise ArgusError("pts must be an Nx2 array")
    pts[:, 0] = (pts[:, 0] - prof[1]) / prof[0]       # in place, as the reference
    pts[:, 1] = (pts[:, 1] - prof[2]) / prof[0]
    return pts


def solve_dlt(xyz, uv):
    """11-parameter DLT fit (tools.py:220-280) on the GPU (argus_dlt_fit): returns (L (11 x 1), mean reprojection distance)."""
    if type(xyz) != np.ndarray or type(uv) != np.ndarray:
        raise ArgusError("uv and xyz must be numpy arrays")
    if xyz.ndim != 2 or uv.ndim != 2:
        raise ArgusError("xyz and uv must be 2-d arrays")
    if xyz.shape[0] != uv.shape[0]:
        raise ArgusError("xyz and uv must have the same number of rows")
    if xyz.shape[1] != 3:
        raise ArgusError("xyz must be an Nx3 array")
    if uv.shape[1] != 2:
        raise ArgusError("uv must be an Nx2 array")
    L, _, (n, total, _, _) = _dlt.fit(xyz, uv, both_coordinates=True)        # rows with a NaN in uv are skipped (tools.py:233-240)
    return L.reshape(11, 1), (total / float(n) if n else float("nan"))


def _check_cams(pts, profs, dlt):
    if (int(pts.shape[1] / 2) != len(profs)) or (int(pts.shape[1] / 2) != len(dlt)):
        raise ArgusError("the length of the profile list and DLT coefficients should match the number of cameras present")
    for p in profs:
        if not _is_pinhole(p):
            raise ArgusError("only pinhole distortion profiles are supported by the B200 backend")


def uv_to_xyz(pts, profs, dlt):
    """pts: N x (2 * cameras) pixel coordinates of ONE track (NaN = not seen) -> N x 3."""
    pts = np.asarray(pts, dtype=np.float64)
    _check_cams(pts, profs, dlt)
    return _dlt.triangulate(pts, profs, np.asarray(dlt, dtype=np.float64))


def get_repo_errors(xyzs, pts, prof, dlt):
    """xyzs: N x 3k, pts: N x (2 * cameras * k) -> k x N reprojection errors."""
    if (not hasattr(prof, "__iter__")) or (not hasattr(dlt, "__iter__")):
        raise ArgusError("camera profile and dlt coefficients must be iterables")
    if type(xyzs) != np.ndarray or xyzs.ndim != 2 or (xyzs.shape[1] % 3) != 0:
        raise ArgusError("xyz values must be an N*(3*k) array, where k is the number of tracks")
    for p in prof:
        if not _is_pinhole(p):
            raise ArgusError("only pinhole distortion profiles are supported by the B200 backend")
    return _dlt.reproj_error(xyzs, np.asarray(pts, dtype=np.float64), prof, np.asarray(dlt, dtype=np.float64))


def uv_to_xyz_and_errors(pts, profs, dlt):
    """All tracks at once (extension): pts N x (2 * cameras * k) -> (xyz N x 3k, errors k x N); equals calling uv_to_xyz on
    every track slice and get_repo_errors on the result, with one host<->device round trip."""
    pts = np.asarray(pts, dtype=np.float64)
    return _dlt.reconstruct(pts, profs, np.asarray(dlt, dtype=np.float64))


def bootstrapXYZs(pts, rmses, prof, dlt, bsIter=250, display_progress=False, subframeinterp=True, seed=None, noise=None):
    """95 % confidence intervals of the triangulated points by bootstrapping (tools.py:448-544): same arguments and return
    values (CIs N x 3k, spline weights, tolerances) as the reference.  The ``bsIter`` perturbed re-triangulations per track run
    as ONE kernel (argus_dlt_bootstrap: on-device Philox noise, Welford standard deviation); the optional sub-frame offset
    search (21 candidate offsets per camera, tools.py:455-510) keeps the reference's host logic with the GPU
    triangulation / reprojection calls inside.  ``seed`` / ``noise`` (extension): reproducible device noise, or explicit
    standard-normal numbers of shape (tracks, bsIter, frames, 2 * cameras) in the reference's draw order."""
    from scipy import interpolate
    numcams = len(prof)
    numpts = int(pts.shape[1] / (2 * numcams))
    dlt = np.asarray(dlt, dtype=np.float64)
    if subframeinterp and numcams > 2:
        print('Checking subframe interpolation')
        redormse = False
        step = list(np.linspace(-1, 1, 21))
        for c in range(1, numcams):
            steprmses = np.zeros((21, numpts)) * np.nan
            ocam = 1 if c > 1 else 2
            for k in range(numpts):
                base = k * 2 * numcams
                c1pts = pts[:, base:base + 2]
                cpts_clean = pts[:, base + c * 2:base + c * 2 + 2]
                ocpts = pts[:, base + ocam * 2:base + ocam * 2 + 2]
                fin = np.where(np.isfinite(cpts_clean[:, 0]))[0]
                if not np.any(fin):
                    continue
                sprof = np.vstack([prof[0], prof[ocam], prof[c]])
                sdlt = np.vstack([dlt[0], dlt[ocam], dlt[c]])
                f = interpolate.interp1d(fin, cpts_clean[fin], axis=0, kind='linear', fill_value='extrapolate')
                for i, s in enumerate(step):
                    cpts = cpts_clean.copy()
                    cpts[fin] = f(fin + s)
                    spts = np.hstack([c1pts, ocpts, cpts])
                    spts = np.ascontiguousarray(spts[np.where(np.sum(np.isfinite(spts), axis=1) > 4)[0]])
                    sxyzs, srms = _dlt.reconstruct(spts, sprof, sdlt)
                    with np.errstate(all='ignore'):
                        import warnings
                        with warnings.catch_warnings():
                            warnings.simplefilter('ignore', category=RuntimeWarning)
                            steprmses[i, k] = np.nanmedian(srms)
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter('ignore', category=RuntimeWarning)
                poff = step[np.nanargmin(np.nanmedian(steprmses, axis=1))]
            if poff != 0:
                redormse = True
                print('partial offset of {} frames found for cam {}, remaking pts'.format(poff, c + 1))
                for k in range(numpts):
                    sl = slice(k * 2 * numcams + c * 2, k * 2 * numcams + c * 2 + 2)
                    cpts = pts[:, sl]
                    # NOTE: the reference indexes with the `fin` of the LAST track's clean points here (tools.py:497-498); kept
                    cpts[fin] = interpolate.interp1d(fin, cpts[fin], axis=0, kind='linear', fill_value='extrapolate')(fin + poff)
                    pts[:, sl] = cpts
        if redormse:
            xyzs, err = _dlt.reconstruct(np.ascontiguousarray(pts), prof, dlt)
            rmses = err.T
    if seed is None and noise is None:
        seed = int(np.random.randint(0, 2 ** 31 - 1))
    ret = _dlt.bootstrap_sd(np.ascontiguousarray(pts), np.asarray(rmses, dtype=np.float64)[:, :numpts], prof, dlt, bs_iter=bsIter, noise=noise, seed=seed or 0)
    import warnings
    with warnings.catch_warnings(), np.errstate(all='ignore'):
        warnings.simplefilter('ignore', category=RuntimeWarning)
        weights = 1 / (ret / np.tile(np.nanmin(ret, axis=0), (ret.shape[0], 1)))
        tols = np.nansum(np.multiply(weights, np.power(ret, 2)), axis=0)
    return ret * 1.96, weights, tols
