"""TEST INFRASTRUCTURE - ctypes wrapper of oracle/ba_port.c (the CPU restatement of the reference BA).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "ba_port.c")
_OUT = os.path.join(_HERE, "_build", "libba_port.so")
_lib = None


def build(force=False):
    if not force and os.path.exists(_OUT) and os.path.getmtime(_OUT) >= os.path.getmtime(_SRC):
        return _OUT
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _OUT, _SRC, "-lm"])
    return _OUT


def load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_motstr.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def port_project(p, vmask, rot0):
    lib = load()
    vmask = np.ascontiguousarray(vmask, dtype=np.int8); n, m = vmask.shape
    p = np.ascontiguousarray(p, dtype=np.float64).ravel(); rot0 = np.ascontiguousarray(rot0, dtype=np.float64).ravel()
    hx = np.zeros((int((vmask != 0).sum()), 2))
    lib.orc_project(ctypes.c_int64(n), ctypes.c_int(m), _p(vmask), _p(p), _p(rot0), _p(hx))
    return hx


def port_jacobian(p, vmask, rot0, nccalib=0, ncdist=0):
    lib = load()
    vmask = np.ascontiguousarray(vmask, dtype=np.int8); n, m = vmask.shape
    p = np.ascontiguousarray(p, dtype=np.float64).ravel(); rot0 = np.ascontiguousarray(rot0, dtype=np.float64).ravel()
    nobs = int((vmask != 0).sum())
    A = np.zeros((nobs, 2, 16)); B = np.zeros((nobs, 2, 3))
    lib.orc_jacobian(ctypes.c_int64(n), ctypes.c_int(m), _p(vmask), _p(p), _p(rot0), ctypes.c_int(nccalib), ctypes.c_int(ncdist), _p(A), _p(B))
    return A, B


def port_motstr(p, x, vmask, rot0, nccalib, ncdist, itmax=150, opts=None, ncon=0, mcon=0, verbose=0, compat=True, cam_fix=None):
    """Same contract as oracle.sba_ref.ref_motstr: returns (p_new, info[10], retval)."""
    lib = load()
    vmask = np.ascontiguousarray(vmask, dtype=np.int8); n, m = vmask.shape
    p = np.array(p, dtype=np.float64, copy=True).ravel()
    x = np.ascontiguousarray(x, dtype=np.float64).ravel(); rot0 = np.ascontiguousarray(rot0, dtype=np.float64).ravel()
    if opts is None:
        opts = [1e-3, 1e-12, 1e-12, 1e-12, 0.0]
    opts = np.ascontiguousarray(opts, dtype=np.float64)
    info = np.zeros(10)
    if cam_fix is not None:      # per-camera (nccalib_j, ncdist_j) extension
        nk = np.ascontiguousarray(cam_fix[0], dtype=np.int32); nd = np.ascontiguousarray(cam_fix[1], dtype=np.int32)
        lib.orc_set_camera_fix(_p(nk), _p(nd))
    ret = lib.orc_motstr(ctypes.c_int64(n), ctypes.c_int64(ncon), ctypes.c_int(m), ctypes.c_int(mcon), _p(vmask), _p(p), _p(rot0), _p(x),
                         ctypes.c_int(nccalib), ctypes.c_int(ncdist), ctypes.c_int(itmax), ctypes.c_int(verbose), _p(opts), _p(info),
                         ctypes.c_int(1 if compat else 0))
    if cam_fix is not None:
        lib.orc_set_camera_fix(None, None)
    return p, info, ret


def _blockdiag(which, p, x, vmask, rot0, nccalib, ncdist, itmax, opts, ncon, mcon, verbose, compat):
    lib = load()
    vmask = np.ascontiguousarray(vmask, dtype=np.int8); n, m = vmask.shape
    p = np.array(p, dtype=np.float64, copy=True).ravel()
    x = np.ascontiguousarray(x, dtype=np.float64).ravel(); rot0 = np.ascontiguousarray(rot0, dtype=np.float64).ravel()
    opts = np.ascontiguousarray([1e-3, 1e-12, 1e-12, 1e-12, 0.0] if opts is None else opts, dtype=np.float64)
    info = np.zeros(10)
    lib.orc_blockdiag.restype = ctypes.c_int
    ret = lib.orc_blockdiag(ctypes.c_int(which), ctypes.c_int64(n), ctypes.c_int64(ncon), ctypes.c_int(m), ctypes.c_int(mcon), _p(vmask), _p(p),
                            _p(rot0), _p(x), ctypes.c_int(nccalib), ctypes.c_int(ncdist), ctypes.c_int(itmax), ctypes.c_int(verbose), _p(opts),
                            _p(info), ctypes.c_int(1 if compat else 0))
    return p, info, ret


def port_mot(p, x, vmask, rot0, nccalib, ncdist, itmax=150, opts=None, mcon=0, verbose=0, compat=True):
    """Motion-only LM (sba_mot_levmar_x, sba_levmar.c:1437): p = [16 m | 3 n], only the cameras move.
    Returns (p_new, info[10], retval)."""
    return _blockdiag(0, p, x, vmask, rot0, nccalib, ncdist, itmax, opts, 0, mcon, verbose, compat)


def port_str(p, x, vmask, rot0, itmax=150, opts=None, ncon=0, verbose=0, compat=True):
    """Structure-only LM (sba_str_levmar_x, sba_levmar.c:1987): p = [16 m | 3 n], only the points move."""
    return _blockdiag(1, p, x, vmask, rot0, 0, 0, itmax, opts, ncon, 0, verbose, compat)
