"""Quaternion helpers with the names python-sba's ``sba.quaternions`` exposes to Argus
(call site: /root/reference/argus_gui/triangulate.py:65-66, ``quatFromRotationMatrix(R).asVector()`` -> [w, x, y, z]).
python-sba 1.6.8.2 itself is not part of the reference tree; conventions follow sba-1.6p/demo/eucsbademo.c:1768-1830
(unit norm, scalar part first and non-negative)."""
import numpy as np

__all__ = ["Quaternion", "quatFromRotationMatrix", "quatFromArray", "quatMult"]


class Quaternion(object):
    def __init__(self, re=1.0, i=0.0, j=0.0, k=0.0):
        self.re, self.i, self.j, self.k = float(re), float(i), float(j), float(k)

    def asVector(self):
        return np.array([self.re, self.i, self.j, self.k])

    def norm(self):
        return float(np.linalg.norm(self.asVector()))

    def normalized(self):
        v = self.asVector()
        v = v / np.linalg.norm(v)
        if v[0] < 0:
            v = -v
        return Quaternion(*v)

    def conjugate(self):
        return Quaternion(self.re, -self.i, -self.j, -self.k)

    def asRotationMatrix(self):
        w, x, y, z = self.normalized().asVector()
        return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                         [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                         [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])

    def __mul__(self, o):
        return quatMult(self, o)

    def __repr__(self):
        return "Quaternion(%r, %r, %r, %r)" % (self.re, self.i, self.j, self.k)


def quatMult(a, b):
    """Hamilton product a (x) b (sbaprojs.c:46-52)."""
    a0, a1, a2, a3 = a.asVector(); b0, b1, b2, b3 = b.asVector()
    return Quaternion(a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3,
                      a0 * b1 + b0 * a1 + a2 * b3 - a3 * b2,
                      a0 * b2 + b0 * a2 + b1 * a3 - a1 * b3,
                      a0 * b3 + b0 * a3 + a1 * b2 - b1 * a2)


def quatFromArray(v):
    v = np.asarray(v, dtype=np.float64).ravel()
    return Quaternion(*v[:4])


def quatFromRotationMatrix(R):
    """Unit quaternion (scalar first, w >= 0) of a 3x3 rotation matrix (numerically safe branch selection)."""
    R = np.asarray(R, dtype=np.float64)
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s]
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s]
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s]
    return Quaternion(*q).normalized()
