"""Duck-type of artelib.homogeneousmatrix.HomogeneousMatrix — the return type of
KeyFrameManager.compute_transformation (reference: artelib/homogeneousmatrix.py:16-107, keyframe.py:259).

When the reference's own `artelib` package is importable (drop-in use inside the reference tree) that class is
used, so callers get exactly the type they expect; otherwise this numpy-only class offers the members the
scan-matching callers touch (run_scanmatcher.py:207-208,221,230; loopclosing.py:89,120-122,182):
`.array`, `*`, `.inv()`, `.pos()`, `.Q()`, `.euler()`, `.t2v()`, `.print_nice()`.
"""
import numpy as np


def rot2quaternion(R):
    """[qw, qx, qy, qz] of a rotation matrix; same branch structure as the reference's artelib.tools.rot2quaternion
    (:110-172): positive scalar part from the trace, vector part scaled from the dominant diagonal entry."""
    R = np.asarray(R, dtype=np.float64)[0:3, 0:3]
    s = np.sqrt(max(0.0, np.trace(R) + 1.0)) / 2.0
    k = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    d = int(np.argmax(np.diag(R)))
    if d == 0:
        k1 = np.array([R[0, 0] - R[1, 1] - R[2, 2] + 1, R[1, 0] + R[0, 1], R[2, 0] + R[0, 2]])
    elif d == 1:
        k1 = np.array([R[1, 0] + R[0, 1], R[1, 1] - R[0, 0] - R[2, 2] + 1, R[2, 1] + R[1, 2]])
    else:
        k1 = np.array([R[2, 0] + R[0, 2], R[2, 1] + R[1, 2], R[2, 2] - R[0, 0] - R[1, 1] + 1])
    sgn = 1.0 if k[d] >= 0 else -1.0
    k = k + sgn * k1
    nm = np.linalg.norm(k)
    if nm == 0:
        return np.array([1.0, 0.0, 0.0, 0.0])
    return np.hstack((s, np.sqrt(max(0.0, 1 - s ** 2)) / nm * k))


def _wrap(angles):
    return np.arctan2(np.sin(angles), np.cos(angles))


def rot2euler(R):
    """Both solutions (alpha, beta, gamma) of R = Rx(alpha) Ry(beta) Rz(gamma) (XYZ, mobile axes), wrapped to
    [-pi, pi] - the convention of the reference's artelib/tools.py:241-275 (pinned by tests/golden/se3_helpers.npz).
    sin(beta) = R[0, 2]; within 1e-4 of gimbal lock alpha is fixed to 0 / pi."""
    R = np.asarray(R, dtype=np.float64)
    s = float(np.clip(R[0, 2], -1.0, 1.0))
    b1 = np.arcsin(s)
    if abs(abs(R[0, 2]) - 1.0) > 1e-4:
        sols = []
        for b in (b1, np.pi - b1):
            c = np.sign(np.cos(b))
            sols.append([np.arctan2(-c * R[1, 2], c * R[2, 2]), b, np.arctan2(-c * R[0, 1], c * R[0, 0])])
    else:
        sg = 1.0 if b1 > 0 else -1.0
        g = np.arctan2(sg * R[1, 0], R[1, 1])
        sols = [[0.0, b1, g], [np.pi, sg * np.pi / 2, g - np.pi]]
    return _wrap(np.array(sols[0])), _wrap(np.array(sols[1]))


def euler2rot(abg):
    """R = Rx(alpha) Ry(beta) Rz(gamma), the inverse of rot2euler (reference artelib/tools.py:226-238)."""
    ca, cb, cg = np.cos(abg[0]), np.cos(abg[1]), np.cos(abg[2])
    sa, sb, sg = np.sin(abg[0]), np.sin(abg[1]), np.sin(abg[2])
    Rx = np.array([[1.0, 0.0, 0.0], [0.0, ca, -sa], [0.0, sa, ca]])
    Ry = np.array([[cb, 0.0, sb], [0.0, 1.0, 0.0], [-sb, 0.0, cb]])
    Rz = np.array([[cg, -sg, 0.0], [sg, cg, 0.0], [0.0, 0.0, 1.0]])
    return Rx @ Ry @ Rz


def quaternion2rot(Q):
    """Unit quaternion [qw, qx, qy, qz] -> rotation matrix."""
    w, x, y, z = (float(v) for v in Q)
    return np.array([[1 - 2 * y ** 2 - 2 * z ** 2, 2 * x * y - 2 * z * w, 2 * x * z + 2 * y * w],
                     [2 * x * y + 2 * z * w, 1 - 2 * x ** 2 - 2 * z ** 2, 2 * y * z - 2 * x * w],
                     [2 * x * z - 2 * y * w, 2 * y * z + 2 * x * w, 1 - 2 * x ** 2 - 2 * y ** 2]])


class Euler:
    """Minimal stand-in of artelib.euler.Euler: the angles live in `.abg`."""
    def __init__(self, abg):
        self.abg = np.asarray(abg.abg if isinstance(abg, Euler) else abg, dtype=np.float64)

    def __str__(self):
        return str(self.abg)


class HomogeneousMatrix:
    def __init__(self, *args):
        if len(args) == 0:
            self.array = np.eye(4)
        elif len(args) == 1:
            a = args[0]
            self.array = a.toarray() if isinstance(a, HomogeneousMatrix) else (a if isinstance(a, np.ndarray) else np.array(a))
        elif len(args) == 2:
            # (position, orientation): orientation = XYZ Euler angles as a list / array / object with `.abg`
            position, orientation = args
            abg = np.asarray(getattr(orientation, "abg", orientation), dtype=np.float64).reshape(3)
            self.array = np.eye(4)
            self.array[:3, :3] = euler2rot(abg)
            self.array[:3, 3] = np.asarray(getattr(position, "array", position), dtype=np.float64).reshape(3)
        else:
            raise TypeError("HomogeneousMatrix(), HomogeneousMatrix(array) or HomogeneousMatrix(position, euler)")

    def __str__(self):
        return str(self.array)

    def toarray(self):
        return self.array

    def print_nice(self, precision=3):
        print(np.array_str(self.array, precision=precision, suppress_small=True))

    def inv(self):
        return HomogeneousMatrix(np.linalg.inv(self.array))

    def Q(self):
        return rot2quaternion(self.array)

    def R(self):
        return self.array[0:3, 0:3]

    def pos(self):
        return self.array[0:3, 3]

    def euler(self):
        e1, e2 = rot2euler(self.array)
        return Euler(e1), Euler(e2)

    def __mul__(self, other):
        if isinstance(other, HomogeneousMatrix) or hasattr(other, "array"):
            return HomogeneousMatrix(np.dot(self.array, other.array))
        return NotImplemented

    def __add__(self, other):
        return HomogeneousMatrix(self.array + other.array)

    def __sub__(self, other):
        return HomogeneousMatrix(self.array - other.array)

    def __getitem__(self, item):
        return self.array[item[0], item[1]]

    def t2v(self, n=2):
        if n == 2:
            return np.array([self.array[0, 3], self.array[1, 3], np.arctan2(self.array[1, 0], self.array[0, 0])])
        e = rot2euler(quaternion2rot(self.Q()))[0]      # via the quaternion, like the reference's t2v (homogeneousmatrix.py:95-107)
        return np.array([self.array[0, 3], self.array[1, 3], self.array[2, 3], e[0], e[1], e[2]])


_NO_REFERENCE = set()


def result_type():
    """The class compute_transformation wraps its result in: the reference's own when available.  Only the NEGATIVE answer is
    cached, per sys.path: a failed import is not cached by Python and costs a path search every time (this is called per
    pose), while a successful one lives in sys.modules - whose current entry must be used, not a remembered class object."""
    import sys
    mod = sys.modules.get("artelib.homogeneousmatrix")
    if mod is not None and hasattr(mod, "HomogeneousMatrix"):
        return mod.HomogeneousMatrix
    key = tuple(sys.path)
    if key in _NO_REFERENCE:
        return HomogeneousMatrix
    try:
        from artelib.homogeneousmatrix import HomogeneousMatrix as RefH   # noqa: WPS433 (reference tree on sys.path)
        return RefH
    except Exception:
        _NO_REFERENCE.add(key)
        return HomogeneousMatrix
