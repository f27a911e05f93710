"""CPU: host-side mirrors of the reference interface: the sba-compatible module (packing / unpacking / text files),
triangulate.multiTriangulator against golden output of the reference's own triangulate.py, sharding helpers."""
import os
import numpy as np
import pytest

from argus_gui_b200 import sba, scenes, triangulate, ba

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_multitriangulator_oracle_golden():
    """The CPU restatement of triangulate.py (oracle/tri_port.py) against the output of the reference module itself; the
    CUDA path is checked against both in tests/test_gpu_pair.py."""
    from oracle import tri_port
    g = np.load(os.path.join(ROOT, "tests", "golden", "tri_golden.npz"))
    for tag in ("A", "B"):
        t = tri_port.multiTriangulator(g[tag + "_pts"].copy(), g[tag + "_intr"])
        xyz = t.triangulate()
        ref = g[tag + "_xyz"]
        assert np.array_equal(np.isnan(xyz), np.isnan(ref))
        ok = ~np.isnan(ref)
        assert np.abs(xyz[ok] - ref[ok]).max() <= 1e-9 * np.abs(ref[ok]).max()
        assert np.abs(t.ext - g[tag + "_ext"]).max() < 1e-9
        assert list(t.ext[-1]) == [1, 0, 0, 0, 0, 0, 0]


def test_sba_containers_roundtrip(tmp_path):
    s = scenes.make_ba_scene(4, 60, 3, seed=2)
    n, m = 60, 4
    dylan = np.full((n, 3 + 2 * m), np.nan)
    dylan[:, :3] = s["pts0"]
    uv = dylan[:, 3:].reshape(n, m, 2); uv[s["vmask"] != 0] = s["x"]
    P = sba.Points.fromDylan(dylan)
    assert np.array_equal(P.vmask, s["vmask"]) and np.array_equal(P.X, s["x"]) and np.array_equal(P.B, s["pts0"])
    assert np.array_equal(np.isnan(P.toDylan()), np.isnan(dylan))
    C = sba.Cameras.fromDylan(s["cams0"])
    a, q = C.pack()
    assert np.all(a[:, 10:13] == 0) and np.allclose(np.linalg.norm(q, axis=1), 1) and np.all(q[:, 0] >= 0)
    want = s["cams0"].copy()
    want[:, 10:14] /= np.linalg.norm(want[:, 10:14], axis=1, keepdims=True)
    want[want[:, 10] < 0, 10:14] *= -1            # pack() normalises and makes the scalar part non-negative
    assert np.abs(sba.Cameras.unpack(a, q) - want).max() < 1e-15
    # a non-zero local rotation composes as q = (sqrt(1-|v|^2), v) (x) q0
    a2 = a.copy(); a2[:, 10:13] = [[0.01, -0.02, 0.03]] * m
    c2 = sba.Cameras.unpack(a2, q)
    for j in range(m):
        ql = sba.quaternions.Quaternion(np.sqrt(1 - 0.0014), 0.01, -0.02, 0.03)
        want = (ql * sba.quaternions.quatFromArray(q[j])).normalized().asVector()
        assert np.abs(c2[j, 10:14] - want).max() < 1e-14
    # text files are np.loadtxt-readable and lossless (sbaDriver.py:299-305 re-reads them)
    C.toTxt(str(tmp_path / "c.txt")); P.toTxt(str(tmp_path / "p.txt"))
    assert np.array_equal(np.loadtxt(str(tmp_path / "c.txt")), s["cams0"])
    assert np.array_equal(np.loadtxt(str(tmp_path / "p.txt")), s["pts0"])
    o = sba.Options.fromInput(C, P)
    o.camera = sba.OPTS_CAMS; o.nccalib = sba.OPTS_FIX5_INTR; o.ncdist = sba.OPTS_FIX4_DIST
    assert (sba.OPTS_FIX5_INTR, sba.OPTS_FIX4_INTR, sba.OPTS_FIX2_INTR) == (5, 4, 2)
    assert (sba.OPTS_FIX5_DIST, sba.OPTS_FIX4_DIST, sba.OPTS_FIX3_DIST, sba.OPTS_FIX0_DIST) == (5, 4, 3, 0)


def test_quaternion_from_rotation():
    rng = np.random.default_rng(0)
    for _ in range(50):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        if q[0] < 0:
            q = -q
        R = scenes.quat_to_rot(q)
        assert np.abs(sba.quaternions.quatFromRotationMatrix(R).asVector() - q).max() < 1e-12
        assert np.abs(sba.quaternions.quatFromArray(q).asRotationMatrix() - R).max() < 1e-12


def test_shard_points():
    n = 1001
    cover = []
    for r in range(4):
        lo, hi = ba.shard_points(n, 4, r)
        cover += list(range(lo, hi))
    assert cover == list(range(n))
    nobs = np.ones(n, dtype=np.int64); nobs[:100] = 50
    cuts = [ba.shard_points(n, 4, r, nobs) for r in range(4)]
    assert cuts[0][0] == 0 and cuts[-1][1] == n and all(cuts[r][1] == cuts[r + 1][0] for r in range(3))
    loads = [nobs[lo:hi].sum() for lo, hi in cuts]
    assert max(loads) - min(loads) <= 60
