"""CPU: the product's camera model (argus_gui_b200/csrc/camera_model.cuh, compiled for the host) against the golden
projection / Jacobian of the reference and a CHKDER-style finite-difference check (sba_chkjac.c:91-211)."""
import ctypes
import os
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_lib():
    out = os.path.join(ROOT, "tests", "_build", "libcamera_model_host.so")
    src = os.path.join(ROOT, "tests", "host", "camera_model_host.cpp")
    hdr = os.path.join(ROOT, "argus_gui_b200", "csrc", "camera_model.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-x", "c++", "-o", out, src])
    return ctypes.CDLL(out)


def _dp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _jac(lib, p, rot0, vmask, nk, nd):
    m = vmask.shape[1]
    pi, ci = np.nonzero(vmask); pi = pi.astype(np.int32); ci = ci.astype(np.int32)
    nobs = len(pi)
    hx = np.zeros((nobs, 2)); A = np.zeros((nobs, 2, 16)); B = np.zeros((nobs, 2, 3))
    lib.host_jacobian(_dp(p), _dp(rot0), m, ctypes.c_int64(nobs), _dp(pi), _dp(ci), nk, nd, _dp(hx), _dp(A), _dp(B))
    return hx, A, B


@pytest.mark.parametrize("nk,nd", [(0, 0), (2, 3), (5, 4)])
def test_against_golden(host_lib, golden_ba, nk, nd):
    g = golden_ba
    p = np.ascontiguousarray(g["G1_p0r"]); rot0 = np.ascontiguousarray(g["G1_rot0"])
    hx, A, B = _jac(host_lib, p, rot0, g["G1_vmask"], nk, nd)
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(hx, g["G1_hx"]) < 1e-13                     # P0
    assert rel(A, g["G1_A_%d%d" % (nk, nd)]) < 1e-11       # P1
    assert rel(B, g["G1_B_%d%d" % (nk, nd)]) < 1e-11


def test_finite_differences(host_lib, golden_ba):
    g = golden_ba
    p = np.ascontiguousarray(g["G1_p0r"]); rot0 = np.ascontiguousarray(g["G1_rot0"]); vm = g["G1_vmask"]
    m = vm.shape[1]
    hx, A, B = _jac(host_lib, p, rot0, vm, 0, 0)
    pi, ci = np.nonzero(vm)
    for col in range(16):
        h = 1e-6 * max(1.0, abs(p[col]))
        pp = p.copy(); pm = p.copy()
        for j in range(m):
            pp[16 * j + col] += h; pm[16 * j + col] -= h
        fd = (_jac(host_lib, pp, rot0, vm, 0, 0)[0] - _jac(host_lib, pm, rot0, vm, 0, 0)[0]) / (2 * h)
        scale = np.abs(A[:, :, col]).max() + 1e-12
        assert np.abs(fd - A[:, :, col]).max() / scale < 1e-6, col
