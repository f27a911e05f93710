"""Shared test helpers: readers for the reference's own demo data sets (sba-1.6p/demo/README.txt:9-82 formats),
comparison utilities."""
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "oracle", "_ref", "demo")

# final mean squared reprojection errors published in the reference (sba-1.6p/demo/README.txt:131-206)
# name -> (cams file, pts file, nccalib, ncdist, initial error, final error, digits)
README_KNOWN = {
    "7cams": ("7cams.txt", "7pts.txt", 5, 5, 19.0947, 0.675396, 6),
    "9cams": ("9cams.txt", "9pts.txt", 5, 5, 8.17604, 0.619559, 6),
    "54cams": ("54cams.txt", "54pts.txt", 5, 5, 2.14707, 0.176473, 6),
    "54camsvarK": ("54camsvarK.txt", "54pts.txt", 2, 5, 2.14707, 0.134928, 6),
    "54camsvarKD": ("54camsvarKD.txt", "54pts.txt", 2, 3, 2.14707, 0.128779, 6),
}


def have_demo():
    return os.path.exists(os.path.join(DEMO, "54pts.txt"))


def _rows(path):
    rows = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            rows.append([float(t) for t in line.split()])
    return rows


def load_demo(cams_file, pts_file):
    """-> cams17 (m x 17), pts (n x 3), vmask (n x m uint8), x (nobs x 2)."""
    cr = _rows(os.path.join(DEMO, cams_file))
    m = len(cr)
    ncol = len(cr[0])
    cams = np.zeros((m, 17))
    if ncol == 7:       # quaternion + translation, intrinsics from calib.txt
        K = np.array(_rows(os.path.join(DEMO, "calib.txt")))
        cams[:, 0] = K[0, 0]; cams[:, 1] = K[0, 2]; cams[:, 2] = K[1, 2]; cams[:, 3] = K[1, 1] / K[0, 0]; cams[:, 4] = K[0, 1]
        cams[:, 10:17] = np.array(cr)
    elif ncol == 12:    # 5 intrinsics + quaternion + translation
        a = np.array(cr); cams[:, 0:5] = a[:, 0:5]; cams[:, 10:17] = a[:, 5:12]
    elif ncol == 17:
        cams[:] = np.array(cr)
    else:
        raise ValueError("unknown camera file format")
    pr = _rows(os.path.join(DEMO, pts_file))
    n = len(pr)
    pts = np.zeros((n, 3)); vmask = np.zeros((n, m), dtype=np.uint8); xs = []
    for i, r in enumerate(pr):
        pts[i] = r[0:3]
        nf = int(r[3])
        obs = sorted((int(r[4 + 3 * t]), r[5 + 3 * t], r[6 + 3 * t]) for t in range(nf))
        for (j, u, v) in obs:
            vmask[i, j] = 1; xs.append((u, v))
    return cams, pts, vmask, np.array(xs)


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def rel_blocks(p_a, p_b, m):
    """max over {camera intrinsics, distortion, rotation, translation, points} of the block-relative difference."""
    ca, cb = p_a[:16 * m].reshape(m, 16), p_b[:16 * m].reshape(m, 16)
    worst = 0.0
    for sl in (slice(0, 5), slice(5, 10), slice(10, 13), slice(13, 16)):
        if np.abs(cb[:, sl]).max() > 0:
            worst = max(worst, rel(ca[:, sl], cb[:, sl]))
    return max(worst, rel(p_a[16 * m:], p_b[16 * m:]))
