"""Drop-in for the DLT functions of /root/reference/argus_gui/tools.py (same names, arguments, return shapes, errors);
the per-frame Python loops of the reference are replaced by the batched CUDA kernels behind the C ABI.

    uv_to_xyz(pts, profs, dlt)            tools.py:296-337   -> argus_dlt_triangulate
    get_repo_errors(xyzs, pts, prof, dlt) tools.py:360-423   -> argus_dlt_reproj_error
    undistort_pts(pts, prof)              tools.py:129-148   -> argus_undistort_points   (pinhole profiles)
    redistort_pts(pts, prof)              tools.py:160-183   (host numpy: only used for the handful of reported outliers)
    solve_dlt(xyz, uv)                    tools.py:220-280   -> argus_dlt_fit
    reconstruct_uv(L, xyz), dlt_inverse(L, xyz), normalize(pts, prof)   (host numpy, not hot)

Only pinhole profiles (list / ndarray) are supported; the omnidirectional model objects of the external ``argus.ocam``
package are out of scope."""
import numpy as np

from . import dlt as _dlt

__all__ = ["ArgusError", "dlt_inverse", "reconstruct_uv", "undistort_pts", "redistort_pts", "normalize", "solve_dlt", "uv_to_xyz",
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


def redistort_pts(pts, prof):
    """Nx2 undistorted pixel coordinates -> Nx2 distorted pixel coordinates: the forward camera model applied to the normalised
    coordinates (the reference projects (x, y, 1) with cv2.projectPoints, identity pose; tools.py:160-183).  ``pts`` is
    normalised in place, as the reference does."""
    if not _is_pinhole(prof):
        raise ArgusError("only pinhole distortion profiles are supported by the B200 backend")
    prof = np.asarray(prof, dtype=np.float64)
    if len(prof) != 8:
        raise ArgusError('pinhole distortion profile must contain exactly 8 coefficients')
    q = normalize(pts, prof)
    x, y = q[:, 0], q[:, 1]
    k1, k2, p1, p2, k3 = prof[3:8]
    r2 = x * x + y * y
    rad = 1.0 + r2 * (k1 + r2 * (k2 + r2 * k3))
    xd = x * rad + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x)
    yd = y * rad + p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y
    return np.stack([prof[0] * xd + prof[1], prof[0] * yd + prof[2]], axis=1)


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


_SUBFRAME_STEPS = np.linspace(-1, 1, 21)


def _subframe_offsets(pts, prof, dlt, numcams, numpts):
    """Sub-frame synchronisation search of bootstrapXYZs (tools.py:455-510): for every camera but the first, the track of
    that camera is re-sampled at 21 time offsets in [-1, 1] frames (linear interpolation), re-triangulated together with the
    first camera and one other camera, and the offset with the smallest median (over tracks) of the median reprojection
    error wins; a non-zero winner is written back into ``pts`` (in place).  All 21 x tracks candidate point sets of a camera
    go through ONE fused GPU triangulation + reprojection call.  Returns True when ``pts`` changed.

    Two details of the reference are kept because its outputs depend on them: a track whose finite rows are exactly [0] is
    skipped (``np.any(fin)``), and the write-back of every track uses the finite rows of the LAST track (tools.py:497-498)."""
    from scipy import interpolate
    nstep = len(_SUBFRAME_STEPS)
    col = lambda k, cam: slice(2 * (numcams * k + cam), 2 * (numcams * k + cam) + 2)
    # values of a track (given on the frames `rows`) at rows + shift; `shift` may be a column of offsets -> (offsets, rows, 2)
    resample = lambda rows, vals, shift: interpolate.interp1d(rows, vals, axis=0, kind='linear', fill_value='extrapolate')(rows + shift)
    changed = False
    for c in range(1, numcams):
        partner = 1 if c > 1 else 2
        trio_prof = np.vstack([prof[0], prof[partner], prof[c]])
        trio_dlt = np.vstack([dlt[0], dlt[partner], dlt[c]])
        blocks, sizes, rows_last = [], np.zeros((nstep, numpts), dtype=np.int64), None
        for k in range(numpts):
            track_c = pts[:, col(k, c)]
            rows_last = np.where(np.isfinite(track_c[:, 0]))[0]
            if not np.any(rows_last):
                continue
            cand = np.repeat(np.hstack([pts[:, col(k, 0)], pts[:, col(k, partner)], track_c])[None], nstep, axis=0)   # steps x frames x 6
            cand[:, rows_last, 4:6] = resample(rows_last, track_c[rows_last], _SUBFRAME_STEPS[:, None])
            usable = np.isfinite(cand).sum(axis=2) > 4
            sizes[:, k] = usable.sum(axis=1)
            blocks.append((k, cand, usable))
        med = np.full((nstep, numpts), np.nan)
        if blocks:
            stacked = np.ascontiguousarray(np.vstack([cand[i][usable[i]] for (_, cand, usable) in blocks for i in range(nstep)]))
            err = _dlt.reconstruct(stacked, trio_prof, trio_dlt)[1].ravel() if stacked.shape[0] else np.zeros(0)
            at = 0
            import warnings
            with warnings.catch_warnings(), np.errstate(all='ignore'):
                warnings.simplefilter('ignore', category=RuntimeWarning)
                for (k, _, _) in blocks:
                    for i in range(nstep):
                        med[i, k] = np.nanmedian(err[at:at + sizes[i, k]])
                        at += sizes[i, k]
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', category=RuntimeWarning)
            best = _SUBFRAME_STEPS[np.nanargmin(np.nanmedian(med, axis=1))]
        if best != 0:
            changed = True
            print('partial offset of {} frames found for cam {}, remaking pts'.format(best, c + 1))
            for k in range(numpts):
                track_c = pts[:, col(k, c)]
                track_c[rows_last] = resample(rows_last, track_c[rows_last], best)
                pts[:, col(k, c)] = track_c
    return changed


def bootstrapXYZs(pts, rmses, prof, dlt, bsIter=250, display_progress=False, subframeinterp=True, seed=None, noise=None):
    """95 % confidence intervals of the triangulated points by bootstrapping (tools.py:448-544): same arguments and return
    values (CIs N x 3k, spline weights, tolerances) as the reference.  The ``bsIter`` perturbed re-triangulations per track run
    as ONE kernel (argus_dlt_bootstrap: on-device Philox noise, Welford standard deviation); the optional sub-frame offset
    search (21 candidate offsets per camera, tools.py:455-510) is one fused GPU call per camera (``_subframe_offsets``);
    like the reference it shifts ``pts`` in place when it finds an offset.  ``seed`` / ``noise`` (extension): reproducible device noise, or explicit
    standard-normal numbers of shape (tracks, bsIter, frames, 2 * cameras) in the reference's draw order."""
    from scipy import interpolate
    numcams = len(prof)
    numpts = int(pts.shape[1] / (2 * numcams))
    dlt = np.asarray(dlt, dtype=np.float64)
    if subframeinterp and numcams > 2 and _subframe_offsets(pts, prof, dlt, numcams, numpts):
        rmses = _dlt.reconstruct(np.ascontiguousarray(pts), prof, dlt)[1].T
    if seed is None and noise is None:
        seed = int(np.random.randint(0, 2 ** 31 - 1))
    ret = _dlt.bootstrap_sd(np.ascontiguousarray(pts), np.asarray(rmses, dtype=np.float64)[:, :numpts], prof, dlt, bs_iter=bsIter, noise=noise, seed=seed or 0)
    import warnings
    with warnings.catch_warnings(), np.errstate(all='ignore'):
        warnings.simplefilter('ignore', category=RuntimeWarning)
        weights = 1 / (ret / np.tile(np.nanmin(ret, axis=0), (ret.shape[0], 1)))
        tols = np.nansum(np.multiply(weights, np.power(ret, 2)), axis=0)
    return ret * 1.96, weights, tols
