"""Synthetic scenes for the BA and DLT hot paths.

Everything is generated with ``numpy.random.default_rng(seed)`` so the same arrays can be handed to the
CUDA path and to the CPU checkers used by the tests.  The camera model is the one of SURVEY.md
appendix A (sba-1.6p/libsbaprojs/imgproj.c:1206-1270): pinhole + Bouguet distortion,
camera row = [f, cx, cy, AR, s, r2, r4, t1, t2, r6, qw, qx, qy, qz, tx, ty, tz].
"""
import numpy as np

__all__ = ["quat_to_rot", "rot_to_quat", "project_points", "make_ring_cameras", "make_ba_scene", "make_ba_wand_scene",
           "make_dlt_scene", "cameras_to_dlt", "pack_ba_problem", "dense_to_sparse"]


def quat_to_rot(q):
    """Rotation matrix of a unit quaternion (scalar first), batched over leading dims."""
    q = np.asarray(q, dtype=np.float64)
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = w * w + x * x - y * y - z * z
    R[..., 0, 1] = 2 * (x * y - w * z)
    R[..., 0, 2] = 2 * (x * z + w * y)
    R[..., 1, 0] = 2 * (x * y + w * z)
    R[..., 1, 1] = w * w - x * x + y * y - z * z
    R[..., 1, 2] = 2 * (y * z - w * x)
    R[..., 2, 0] = 2 * (x * z - w * y)
    R[..., 2, 1] = 2 * (y * z + w * x)
    R[..., 2, 2] = w * w - x * x - y * y + z * z
    return R


def rot_to_quat(R):
    """Unit quaternion (scalar first, w >= 0) of one rotation matrix."""
    R = np.asarray(R, dtype=np.float64)
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s])
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = np.array([(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s])
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = np.array([(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s])
    q /= np.linalg.norm(q)
    if q[0] < 0:
        q = -q
    return q


def project_points(cams, pts, cam_idx):
    """Project ``pts`` (k x 3) through camera rows ``cams[cam_idx]`` (k,) -> (k x 2) pixels."""
    c = cams[cam_idx]
    R = quat_to_rot(c[:, 10:14])
    P = np.einsum("kij,kj->ki", R, pts) + c[:, 14:17]
    x = P[:, 0] / P[:, 2]
    y = P[:, 1] / P[:, 2]
    r2 = x * x + y * y
    k1, k2, p1, p2, k3 = c[:, 5], c[:, 6], c[:, 7], c[:, 8], c[:, 9]
    rho = 1 + r2 * (k1 + r2 * (k2 + r2 * k3))
    xd = rho * x + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = rho * y + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    u = c[:, 0] * xd + c[:, 4] * yd + c[:, 1]
    v = c[:, 0] * c[:, 3] * yd + c[:, 2]
    return np.stack([u, v], axis=1)


def make_ring_cameras(m, radius=8.0, height_amp=1.0, dist=(-0.05, 0.01, 0.001, -0.001, 0.0),
                      f=1000.0, cx=959.5, cy=539.5, arc=2 * np.pi):
    """m cameras on a ring (sinusoidal height) looking at the origin; rows of 17."""
    cams = np.zeros((m, 17))
    for j in range(m):
        ang = arc * j / m
        C = np.array([radius * np.cos(ang), radius * np.sin(ang), height_amp * np.sin(2 * ang + 0.3)])
        zc = -C / np.linalg.norm(C)
        up = np.array([0.0, 0.0, 1.0])
        xc = np.cross(up, zc)
        xc /= np.linalg.norm(xc)
        yc = np.cross(zc, xc)
        R = np.stack([xc, yc, zc])           # world -> camera
        cams[j, 0:5] = [f, cx, cy, 1.0, 0.0]
        cams[j, 5:10] = dist
        cams[j, 10:14] = rot_to_quat(R)
        cams[j, 14:17] = -R @ C
    return cams


def make_ba_scene(m, n, views, seed, noise=0.5, radius=8.0, perturb=True, outlier_frac=0.0,
                  dist=(-0.05, 0.01, 0.001, -0.001, 0.0), point_sigma=1.0, pert_scale=1.0, cam_seed=None, pts=None):
    """Synthetic BA problem (SURVEY.md 8d, configs C3/C4/C5 by argument).

    Returns a dict with the truth, the perturbed start, and the *sparse* problem in the layout the
    C-ABI takes: ``vmask`` (n x m uint8), ``x`` (nobs x 2, point-major, cameras ascending),
    ``cams0`` (m x 17), ``pts0`` (n x 3).
    """
    rng = np.random.default_rng(seed)
    cams = make_ring_cameras(m, radius=radius, dist=dist)
    pts = rng.normal(0.0, point_sigma, size=(n, 3)) if pts is None else np.array(pts, dtype=np.float64).reshape(n, 3)
    views = min(views, m)
    if views == m:
        vmask = np.ones((n, m), dtype=np.uint8)
    else:
        # each point seen by `views` distinct random cameras
        keys = rng.random((n, m))
        kth = np.partition(keys, views - 1, axis=1)[:, views - 1:views]
        vmask = (keys <= kth).astype(np.uint8)
    pi, ci = np.nonzero(vmask)
    x = project_points(cams, pts[pi], ci)
    x += rng.normal(0.0, noise, size=x.shape)
    if outlier_frac > 0:
        nout = int(outlier_frac * n)
        opts_ = rng.choice(n, size=nout, replace=False)
        rowptr = np.zeros(n + 1, dtype=np.int64)
        rowptr[1:] = np.cumsum(vmask.sum(axis=1))
        k = rowptr[opts_] + rng.integers(0, views, size=nout)
        x[k, 0] = rng.uniform(0, 1920, size=nout)
        x[k, 1] = rng.uniform(0, 1080, size=nout)
    cams0 = cams.copy()
    pts0 = pts.copy()
    # the camera perturbation can be seeded separately so that every rank of a sharded run (points seeded per rank)
    # starts from the same cameras
    crng = rng if cam_seed is None else np.random.default_rng(cam_seed)
    if perturb:
        s = pert_scale
        cams0[:, 0] *= 1.0 + 0.01 * s
        cams0[:, 5:10] *= 1.0 - 0.1 * s
        dq = np.concatenate([np.ones((m, 1)), crng.normal(0, 0.0025 * s, size=(m, 3))], axis=1)
        dq /= np.linalg.norm(dq, axis=1, keepdims=True)
        q = cams0[:, 10:14]
        # q <- dq (x) q  (Hamilton product)
        w1, x1, y1, z1 = dq.T
        w2, x2, y2, z2 = q.T
        cams0[:, 10:14] = np.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
                                    w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                                    w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                                    w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], axis=1)
        cams0[:, 14:17] += crng.uniform(-0.02 * s, 0.02 * s, size=(m, 3))
        pts0 += rng.uniform(-0.02 * s, 0.02 * s, size=(n, 3))
    return {"m": m, "n": n, "cams_true": cams, "pts_true": pts, "cams0": cams0, "pts0": pts0,
            "vmask": vmask, "x": x, "nobs": int(x.shape[0])}


def make_ba_wand_scene(m, npairs, nextra, views, seed, wand=0.5, point_sigma=1.0, **kw):
    """A BA scene whose first 2 * npairs points are the two ends of ``npairs`` wand positions (ends 1 first, then ends 2 - the
    order in which Argus' driver stacks its paired points, argus_gui/sbaDriver.py:183-197), followed by ``nextra`` free
    points.  Adds ``ia``, ``ib`` (the point indices of every pair) and ``wand`` to the dict of ``make_ba_scene``."""
    rng = np.random.default_rng(seed + 7919)
    centre = rng.normal(0.0, point_sigma, size=(npairs, 3))
    d = rng.normal(size=(npairs, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = np.vstack([centre - 0.5 * wand * d, centre + 0.5 * wand * d, rng.normal(0.0, point_sigma, size=(nextra, 3))])
    s = make_ba_scene(m, 2 * npairs + nextra, views, seed, point_sigma=point_sigma, pts=pts, **kw)
    s["ia"] = np.arange(npairs, dtype=np.int64); s["ib"] = np.arange(npairs, 2 * npairs, dtype=np.int64); s["wand"] = wand
    return s


def pack_ba_problem(cams17, pts):
    """Camera rows (m x 17) + points -> (p, rot0) as sba sees them (readparams.c:161-224,
    eucsbademo.c:1503-1511): p = [16 per camera with the local rotation zeroed | 3 per point],
    rot0 = unit quaternions with non-negative scalar part."""
    cams17 = np.asarray(cams17, dtype=np.float64)
    m = cams17.shape[0]
    q = cams17[:, 10:14].copy()
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 0] < 0] *= -1
    a = np.zeros((m, 16))
    a[:, 0:10] = cams17[:, 0:10]
    a[:, 13:16] = cams17[:, 14:17]
    p = np.concatenate([a.ravel(), np.asarray(pts, dtype=np.float64).ravel()])
    return p, q.ravel().copy()


def dense_to_sparse(dylan, m):
    """``[X Y Z u1 v1 ... um vm]`` rows with NaN = invisible -> (pts, vmask, x)."""
    dylan = np.asarray(dylan, dtype=np.float64)
    uv = dylan[:, 3:3 + 2 * m].reshape(-1, m, 2)
    vis = ~np.isnan(uv).any(axis=2)
    x = uv[vis]
    return dylan[:, :3].copy(), vis.astype(np.uint8), np.ascontiguousarray(x)


def cameras_to_dlt(cams17):
    """11 DLT coefficients per camera from the *undistorted* pinhole part: P = K [R|t], L = P / P[2,3]."""
    cams17 = np.asarray(cams17, dtype=np.float64)
    out = np.zeros((cams17.shape[0], 11))
    for j, c in enumerate(cams17):
        K = np.array([[c[0], c[4], c[1]], [0.0, c[0] * c[3], c[2]], [0.0, 0.0, 1.0]])
        R = quat_to_rot(c[10:14])
        P = K @ np.concatenate([R, c[14:17, None]], axis=1)
        P = P / P[2, 3]
        out[j] = P.ravel()[:11]
    return out


def make_dlt_scene(nframes, m=4, ntracks=10, seed=2, nan_frac=0.1, noise=0.5, radius=5.0):
    """Config C2: ``pts`` (nframes x 2*m*ntracks, track-major blocks of 2m columns), profiles (m x 8)
    ``[f, cx, cy, k1, k2, p1, p2, k3]`` and DLT coefficients (m x 11)."""
    rng = np.random.default_rng(seed)
    dist = (-0.05, 0.01, 0.001, -0.001, 0.0)
    cams = make_ring_cameras(m, radius=radius, height_amp=0.5, dist=dist)
    dlt = cameras_to_dlt(cams)
    prof = np.stack([np.array([c[0], c[1], c[2], c[5], c[6], c[7], c[8], c[9]]) for c in cams])
    pts = np.empty((nframes, 2 * m * ntracks))
    xyz_true = rng.uniform(-1.0, 1.0, size=(nframes, ntracks, 3))
    for t in range(ntracks):
        for j in range(m):
            uv = project_points(cams, xyz_true[:, t, :], np.full(nframes, j))
            uv += rng.normal(0, noise, size=uv.shape)
            drop = rng.random(nframes) < nan_frac
            uv[drop] = np.nan
            pts[:, t * 2 * m + 2 * j: t * 2 * m + 2 * j + 2] = uv
    return {"pts": pts, "prof": prof, "dlt": dlt, "xyz_true": xyz_true.reshape(nframes, 3 * ntracks),
            "cams": cams, "m": m, "ntracks": ntracks}


def make_wand_scene(nframes=2000, nbackground=100, m=3, seed=1, wand=0.5, noise=0.5, dropout=0.05, radius=5.0):
    """Config C1: m cameras on a 120-degree arc looking at the origin, ``nframes`` frames of a ``wand`` m
    wand (random centre in a 2 m cube, random orientation) -> paired points (nframes x 4m: [end-1 uv per camera | end-2 uv
    per camera]), ``nbackground`` points in a 3 m cube -> unpaired (nbackground x 2m), pinhole + r2 = -0.05 cameras
    (m x 10: f, cx, cy, AR, s, r2, r4, t1, t2, r6), pixel noise, random drop-outs."""
    rng = np.random.default_rng(seed)
    cams = make_ring_cameras(m, radius=radius, height_amp=0.3, dist=(-0.05, 0.0, 0.0, 0.0, 0.0), arc=2 * np.pi / 3)
    centre = rng.uniform(-1.0, 1.0, size=(nframes, 3))
    d = rng.normal(size=(nframes, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    e1, e2 = centre - 0.5 * wand * d, centre + 0.5 * wand * d
    bg = rng.uniform(-1.5, 1.5, size=(nbackground, 3))

    def proj(X):
        out = np.empty((X.shape[0], 2 * m))
        for j in range(m):
            uv = project_points(cams, X, np.full(X.shape[0], j)) + rng.normal(0, noise, size=(X.shape[0], 2))
            uv[rng.random(X.shape[0]) < dropout] = np.nan
            out[:, 2 * j:2 * j + 2] = uv
        return out
    ppts = np.hstack([proj(e1), proj(e2)])
    uppts = proj(bg)
    return {"ppts": ppts, "uppts": uppts, "cams": cams[:, :10].copy(), "cams_true": cams, "wand": wand, "e1": e1, "e2": e2, "bg": bg}
