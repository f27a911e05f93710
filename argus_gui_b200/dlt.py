"""Python host of the DLT path (C ABI: argus_dlt_*).  Array-level functions; the reference-named wrappers
(uv_to_xyz, get_repo_errors, undistort_pts) live in argus_gui_b200/tools.py."""
import ctypes
import numpy as np
from . import _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(pts, prof, dlt, k=None):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    prof = np.ascontiguousarray(np.asarray([np.asarray(p, dtype=np.float64) for p in prof]), dtype=np.float64)
    dlt = np.ascontiguousarray(dlt, dtype=np.float64)
    m = prof.shape[0]
    if prof.ndim != 2 or prof.shape[1] < 8:
        raise ValueError("pinhole profiles must have 8 values [f, cx, cy, k1, k2, p1, p2, k3]")
    if prof.shape[1] > 8:       # tools.undistort_pts uses prof[0:3] and prof[-5:]
        prof = np.ascontiguousarray(np.hstack([prof[:, :3], prof[:, -5:]]))
    if dlt.shape != (m, 11):
        raise ValueError("DLT coefficients must be (cameras x 11)")
    if pts.ndim != 2 or pts.shape[1] % (2 * m) != 0:
        raise ValueError("pts must be N x (2 * cameras * tracks)")
    kk = pts.shape[1] // (2 * m)
    if k is not None and k != kk:
        raise ValueError("track count mismatch")
    return pts, prof, dlt, m, kk


def triangulate(pts, prof, dlt, textbook=False):
    """pts N x (2 m k) -> xyz N x 3k (argus_dlt_triangulate)."""
    lib = _lib.load()
    pts, prof, dlt, m, k = _prep(pts, prof, dlt)
    N = pts.shape[0]
    xyz = np.empty((N, 3 * k))
    _lib.check(lib.argus_dlt_triangulate(_ptr(pts), N, m, k, _ptr(prof), _ptr(dlt), _ptr(xyz), 1 if textbook else 0), "argus_dlt_triangulate")
    return xyz


def reproj_error(xyz, pts, prof, dlt):
    """xyz N x 3k, pts N x 2mk -> err k x N (argus_dlt_reproj_error)."""
    lib = _lib.load()
    pts, prof, dlt, m, k = _prep(pts, prof, dlt)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    N = pts.shape[0]
    if xyz.shape != (N, 3 * k):
        raise ValueError("xyz must be N x 3k")
    err = np.empty((k, N))
    _lib.check(lib.argus_dlt_reproj_error(_ptr(xyz), _ptr(pts), N, m, k, _ptr(prof), _ptr(dlt), _ptr(err)), "argus_dlt_reproj_error")
    return err


def reconstruct(pts, prof, dlt, textbook=False, out=None):
    """Both in one device round trip: -> (xyz N x 3k, err k x N).  ``out=(xyz, err)``: preallocated C-contiguous float64
    results (e.g. pinned memory, so that the download overlaps the upload at full PCIe speed)."""
    lib = _lib.load()
    pts, prof, dlt, m, k = _prep(pts, prof, dlt)
    N = pts.shape[0]
    if out is None:
        xyz = np.empty((N, 3 * k)); err = np.empty((k, N))
    else:
        xyz, err = out
        for a, shp in ((xyz, (N, 3 * k)), (err, (k, N))):
            if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.shape == shp):
                raise ValueError("out must be C-contiguous float64 arrays of shape (N, 3k) and (k, N)")
    _lib.check(lib.argus_dlt_reconstruct(_ptr(pts), N, m, k, _ptr(prof), _ptr(dlt), _ptr(xyz), _ptr(err), 1 if textbook else 0),
               "argus_dlt_reconstruct")
    return xyz, err


def undistort(pts, prof8):
    """N x 2 pixels -> N x 2 float32, faithful to cv2.undistortPoints(src f32, K, dist, P=K)."""
    lib = _lib.load()
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 2)
    prof8 = np.asarray(prof8, dtype=np.float64)
    prof8 = np.ascontiguousarray(np.concatenate([prof8[:3], prof8[-5:]]))
    out = np.empty((pts.shape[0], 2), dtype=np.float32)
    _lib.check(lib.argus_undistort_points(_ptr(pts), pts.shape[0], _ptr(prof8), _ptr(out)), "argus_undistort_points")
    return out


def bootstrap_sd(pts, rmses, prof, dlt, bs_iter=250, noise=None, seed=0):
    """Sample standard deviation (ddof = 1) of ``bs_iter`` perturbed re-triangulations per (frame, track):
    pts N x 2mk, rmses N x k -> N x 3k (argus_dlt_bootstrap).  ``noise``: optional explicit standard-normal numbers
    of shape (k, bs_iter, N, 2m)."""
    lib = _lib.load()
    pts, prof, dlt, m, k = _prep(pts, prof, dlt)
    N = pts.shape[0]
    rmses = np.ascontiguousarray(rmses, dtype=np.float64).reshape(N, k)
    if noise is not None:
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        if noise.shape != (k, bs_iter, N, 2 * m):
            raise ValueError("noise must have shape (tracks, bs_iter, frames, 2 * cameras)")
    sd = np.empty((N, 3 * k))
    rc = lib.argus_dlt_bootstrap(_ptr(pts), N, m, k, _ptr(prof), _ptr(dlt), _ptr(rmses), int(bs_iter),
                                 _ptr(noise) if noise is not None else None, ctypes.c_uint64(int(seed)), _ptr(sd))
    _lib.check(rc, "argus_dlt_bootstrap")
    return sd


def fit(xyz, uv, both_coordinates=False):
    """Least-squares 11-parameter DLT fit of one camera (argus_dlt_fit): xyz N x 3, uv N x 2 (NaN = not seen) ->
    (L (11,), errors (N,; NaN where skipped), (n_used, sum, mean, std))."""
    lib = _lib.load()
    xyz = np.ascontiguousarray(xyz, dtype=np.float64); uv = np.ascontiguousarray(uv, dtype=np.float64)
    if xyz.ndim != 2 or xyz.shape[1] != 3 or uv.shape != (xyz.shape[0], 2):
        raise ValueError("xyz must be N x 3 and uv N x 2")
    N = xyz.shape[0]
    L = np.empty(11); err = np.empty(N); stats = np.empty(4)
    rc = lib.argus_dlt_fit(_ptr(xyz), _ptr(uv), N, 1 if both_coordinates else 0, _ptr(L), _ptr(err), _ptr(stats))
    if rc == _lib.ARGUS_E_SBA:
        raise ArithmeticError("argus_dlt_fit: rank-deficient system (the points seen by this camera do not span 3-D space)")
    _lib.check(rc, "argus_dlt_fit")
    return L, err, (int(stats[0]), float(stats[1]), float(stats[2]), float(stats[3]))
