"""TEST INFRASTRUCTURE - ctypes driver for the reference's own compiled C code (oracle/_ref).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  It is never on the product path.

It drives the *unmodified* sba 1.6 + libsbaprojs binaries that oracle/build_ref.sh compiles from
/root/reference/argus_gui/resources/sba_lib_sources.tar.gz:

* ``ref_motstr``  -> ``sba_motstr_levmar_x``      (sba-1.6p/sba_levmar.c:449-1428, prototype sba.h:111-116)
  with the callbacks ``img_projsKDRTS_x`` / ``img_projsKDRTS_jac_x``
  (sba-1.6p/libsbaprojs/sbaprojs.c:1135-1184, 1193-1250) and ``struct globs_`` (sbaprojs.h:123-156).
* ``ref_mot`` / ``ref_str`` -> ``sba_mot_levmar_x`` / ``sba_str_levmar_x`` (sba_levmar.c:1437, 1987) with the
  ``img_projsKDRT_*`` / ``img_projsKDS_*`` callbacks.
* ``ref_project`` / ``ref_jacobian`` call the two callbacks directly on a CRS index built the way
  sba_levmar.c:640-650 builds it, so the projection and the Jacobian can be compared in isolation.

This is the binding python-sba 1.6.8.2 (absent from the reference tree) would provide; the struct
layouts follow sba-1.6p/libsbaprojs/sbaprojs.py:20-45.
"""
import ctypes
import glob
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")

CNP, PNP, MNP = 16, 3, 2


class Globs(ctypes.Structure):
    _fields_ = [("rot0params", ctypes.POINTER(ctypes.c_double)),
                ("intrcalib", ctypes.POINTER(ctypes.c_double)),
                ("nccalib", ctypes.c_int),
                ("ncdist", ctypes.c_int),
                ("cnp", ctypes.c_int),
                ("pnp", ctypes.c_int),
                ("mnp", ctypes.c_int),
                ("ptparams", ctypes.POINTER(ctypes.c_double)),
                ("camparams", ctypes.POINTER(ctypes.c_double))]


class Crsm(ctypes.Structure):
    _fields_ = [("nr", ctypes.c_int), ("nc", ctypes.c_int), ("nnz", ctypes.c_int),
                ("val", ctypes.POINTER(ctypes.c_int)),
                ("colidx", ctypes.POINTER(ctypes.c_int)),
                ("rowptr", ctypes.POINTER(ctypes.c_int))]


_libs = None


def available():
    return os.path.exists(os.path.join(_REF, "libsba.so")) and os.path.exists(os.path.join(_REF, "libsbaprojs.so"))


def _find_blas_dir():
    p = os.path.join(_REF, "blas_path.txt")
    if os.path.exists(p):
        d = os.path.dirname(open(p).read().strip())
        if os.path.isdir(d):
            return d
    import importlib.util
    spec = importlib.util.find_spec("cv2")
    base = os.path.dirname(os.path.dirname(spec.origin))
    for d in glob.glob(os.path.join(base, "opencv*libs")):
        if glob.glob(os.path.join(d, "libopenblas*.so*")):
            return d
    raise RuntimeError("no OpenBLAS found for the reference libsba.so")


def load():
    """Load libsba.so / libsbaprojs.so (and their OpenBLAS) once."""
    global _libs
    if _libs is not None:
        return _libs
    if not available():
        raise RuntimeError("oracle/_ref is not built; run oracle/build_ref.sh where /root/reference exists")
    # the reference is single-threaded; BLAS threading changes its iteration counts
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    d = _find_blas_dir()
    for pat in ("libquadmath*.so*", "libgfortran*.so*", "libopenblas*.so*"):
        for f in sorted(glob.glob(os.path.join(d, pat))):
            ctypes.CDLL(f, mode=ctypes.RTLD_GLOBAL)
    sba = ctypes.CDLL(os.path.join(_REF, "libsba.so"), mode=ctypes.RTLD_GLOBAL)
    prj = ctypes.CDLL(os.path.join(_REF, "libsbaprojs.so"), mode=ctypes.RTLD_GLOBAL)
    sba.sba_motstr_levmar_x.restype = ctypes.c_int
    sba.sba_motstr_levmar_x.argtypes = [
        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p,
        ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_int,
        ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.c_int,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    # motion-only / structure-only drivers (sba.h:118-130): no pnp (mot) / no cnp (str) argument, one fixed-count argument
    sba.sba_mot_levmar_x.restype = ctypes.c_int
    sba.sba_mot_levmar_x.argtypes = [
        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p,
        ctypes.POINTER(ctypes.c_double), ctypes.c_int,
        ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.c_int,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    sba.sba_str_levmar_x.restype = ctypes.c_int
    sba.sba_str_levmar_x.argtypes = sba.sba_mot_levmar_x.argtypes
    for fn in (prj.img_projsKDRTS_x, prj.img_projsKDRTS_jac_x, prj.img_projsKDRT_x, prj.img_projsKDRT_jac_x,
               prj.img_projsKDS_x, prj.img_projsKDS_jac_x):
        fn.restype = None
        fn.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(Crsm),
                       ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                       ctypes.POINTER(ctypes.c_double), ctypes.c_void_p]
    _libs = (sba, prj)
    return _libs


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _iptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _globs(rot0, nccalib, ncdist):
    g = Globs()
    g.rot0params = _dptr(rot0)
    g.intrcalib = None
    g.nccalib = int(nccalib)
    g.ncdist = int(ncdist)
    g.cnp, g.pnp, g.mnp = CNP, PNP, MNP
    g.ptparams = None
    g.camparams = None
    return g


def _crs(vmask):
    """CRS index as sba_levmar.c:640-650 builds it (val[k]=k, point-major, cameras ascending)."""
    n, m = vmask.shape
    vis = vmask != 0
    nnz = int(vis.sum())
    rowptr = np.zeros(n + 1, dtype=np.int32)
    rowptr[1:] = np.cumsum(vis.sum(axis=1))
    colidx = np.nonzero(vis)[1].astype(np.int32)
    val = np.arange(nnz, dtype=np.int32)
    c = Crsm()
    c.nr, c.nc, c.nnz = n, m, nnz
    c.val, c.colidx, c.rowptr = _iptr(val), _iptr(colidx), _iptr(rowptr)
    return c, (val, colidx, rowptr)


def ref_motstr(p, x, vmask, rot0, nccalib, ncdist, itmax=150, opts=None, ncon=0, mcon=0, verbose=0):
    """Run the reference LM.  ``p`` = [16*m camera params (local rotation in 10:13) | 3*n points] is
    updated on a copy.  Returns (p_new, info[10], retval)."""
    sba, prj = load()
    vmask = np.ascontiguousarray(vmask, dtype=np.int8)
    n, m = vmask.shape
    p = np.array(p, dtype=np.float64, copy=True).ravel()
    x = np.ascontiguousarray(x, dtype=np.float64).ravel()
    rot0 = np.ascontiguousarray(rot0, dtype=np.float64).ravel()
    assert p.size == CNP * m + PNP * n and rot0.size == 4 * m
    if opts is None:
        opts = [1e-3, 1e-12, 1e-12, 1e-12, 0.0]
    opts = np.ascontiguousarray(opts, dtype=np.float64)
    info = np.zeros(10, dtype=np.float64)
    g = _globs(rot0, nccalib, ncdist)
    ret = sba.sba_motstr_levmar_x(
        n, ncon, m, mcon, vmask.ctypes.data_as(ctypes.c_char_p), _dptr(p), CNP, PNP, _dptr(x), None, MNP,
        ctypes.cast(prj.img_projsKDRTS_x, ctypes.c_void_p), ctypes.cast(prj.img_projsKDRTS_jac_x, ctypes.c_void_p),
        ctypes.cast(ctypes.pointer(g), ctypes.c_void_p), int(itmax), int(verbose), _dptr(opts), _dptr(info))
    return p, info, ret


def ref_mot(cams, pts, x, vmask, rot0, nccalib, ncdist, itmax=150, opts=None, mcon=0, verbose=0):
    """Motion-only reference LM: ``sba_mot_levmar_x`` (sba-1.6p/sba_levmar.c:1437-1980) with the callbacks
    ``img_projsKDRT_x`` / ``img_projsKDRT_jac_x`` (libsbaprojs/sbaprojs.c:1261-1370); the fixed points travel in
    ``globs_.ptparams``.  ``cams`` (16 m) is updated on a copy.  Returns (cams_new, info[10], retval)."""
    sba, prj = load()
    vmask = np.ascontiguousarray(vmask, dtype=np.int8)
    n, m = vmask.shape
    cams = np.array(cams, dtype=np.float64, copy=True).ravel()
    pts = np.ascontiguousarray(pts, dtype=np.float64).ravel()
    x = np.ascontiguousarray(x, dtype=np.float64).ravel()
    rot0 = np.ascontiguousarray(rot0, dtype=np.float64).ravel()
    assert cams.size == CNP * m and pts.size == PNP * n and rot0.size == 4 * m
    opts = np.ascontiguousarray([1e-3, 1e-12, 1e-12, 1e-12, 0.0] if opts is None else opts, dtype=np.float64)
    info = np.zeros(10, dtype=np.float64)
    g = _globs(rot0, nccalib, ncdist)
    g.ptparams = _dptr(pts)
    ret = sba.sba_mot_levmar_x(
        n, m, mcon, vmask.ctypes.data_as(ctypes.c_char_p), _dptr(cams), CNP, _dptr(x), None, MNP,
        ctypes.cast(prj.img_projsKDRT_x, ctypes.c_void_p), ctypes.cast(prj.img_projsKDRT_jac_x, ctypes.c_void_p),
        ctypes.cast(ctypes.pointer(g), ctypes.c_void_p), int(itmax), int(verbose), _dptr(opts), _dptr(info))
    return cams, info, ret


def ref_str(cams, pts, x, vmask, itmax=150, opts=None, ncon=0, verbose=0):
    """Structure-only reference LM: ``sba_str_levmar_x`` (sba-1.6p/sba_levmar.c:1987-2540) with the callbacks
    ``img_projsKDS_x`` / ``img_projsKDS_jac_x`` (libsbaprojs/sbaprojs.c:1375-1445).  The fixed cameras travel in
    ``globs_.camparams``; these callbacks take entries 10:13 of a camera row as the vector part of its FULL rotation
    quaternion (there is no rot0 on this path).  ``pts`` (3 n) is updated on a copy."""
    sba, prj = load()
    vmask = np.ascontiguousarray(vmask, dtype=np.int8)
    n, m = vmask.shape
    cams = np.ascontiguousarray(cams, dtype=np.float64).ravel()
    pts = np.array(pts, dtype=np.float64, copy=True).ravel()
    x = np.ascontiguousarray(x, dtype=np.float64).ravel()
    assert cams.size == CNP * m and pts.size == PNP * n
    opts = np.ascontiguousarray([1e-3, 1e-12, 1e-12, 1e-12, 0.0] if opts is None else opts, dtype=np.float64)
    info = np.zeros(10, dtype=np.float64)
    rot0 = np.tile([1.0, 0.0, 0.0, 0.0], m)
    g = _globs(rot0, 0, 0)
    g.camparams = _dptr(cams)
    ret = sba.sba_str_levmar_x(
        n, ncon, m, vmask.ctypes.data_as(ctypes.c_char_p), _dptr(pts), PNP, _dptr(x), None, MNP,
        ctypes.cast(prj.img_projsKDS_x, ctypes.c_void_p), ctypes.cast(prj.img_projsKDS_jac_x, ctypes.c_void_p),
        ctypes.cast(ctypes.pointer(g), ctypes.c_void_p), int(itmax), int(verbose), _dptr(opts), _dptr(info))
    return pts, info, ret


def ref_project(p, vmask, rot0):
    """hx (nobs x 2), point-major / cameras ascending, from img_projsKDRTS_x."""
    _, prj = load()
    vmask = np.ascontiguousarray(vmask, dtype=np.int8)
    n, m = vmask.shape
    p = np.ascontiguousarray(p, dtype=np.float64).ravel().copy()
    rot0 = np.ascontiguousarray(rot0, dtype=np.float64).ravel()
    c, keep = _crs(vmask)
    hx = np.zeros(c.nnz * MNP)
    w1 = np.zeros(max(n, m), dtype=np.int32)
    w2 = np.zeros(max(n, m), dtype=np.int32)
    g = _globs(rot0, 0, 0)
    prj.img_projsKDRTS_x(_dptr(p), ctypes.byref(c), _iptr(w1), _iptr(w2), _dptr(hx),
                         ctypes.cast(ctypes.pointer(g), ctypes.c_void_p))
    return hx.reshape(-1, 2)


def ref_jacobian(p, vmask, rot0, nccalib=0, ncdist=0):
    """(A (nobs x 2 x 16), B (nobs x 2 x 3)) from img_projsKDRTS_jac_x, fixed columns zeroed."""
    _, prj = load()
    vmask = np.ascontiguousarray(vmask, dtype=np.int8)
    n, m = vmask.shape
    p = np.ascontiguousarray(p, dtype=np.float64).ravel().copy()
    rot0 = np.ascontiguousarray(rot0, dtype=np.float64).ravel()
    c, keep = _crs(vmask)
    jac = np.zeros(c.nnz * (MNP * CNP + MNP * PNP))
    w1 = np.zeros(max(n, m), dtype=np.int32)
    w2 = np.zeros(max(n, m), dtype=np.int32)
    g = _globs(rot0, nccalib, ncdist)
    prj.img_projsKDRTS_jac_x(_dptr(p), ctypes.byref(c), _iptr(w1), _iptr(w2), _dptr(jac),
                             ctypes.cast(ctypes.pointer(g), ctypes.c_void_p))
    jac = jac.reshape(c.nnz, MNP * CNP + MNP * PNP)
    return jac[:, :32].reshape(-1, 2, 16).copy(), jac[:, 32:].reshape(-1, 2, 3).copy()
