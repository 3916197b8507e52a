"""Batched SE(2) helpers in numpy float64, rows = samples (x, y, theta).

Restates the algebra of the reference's one-object-per-pose classes
(src/geometry/TwoDimension.py: Rot2 149-300, SE2Pose 303-544) for whole sample arrays: angles are
wrapped to [-pi, pi) whenever a rotation is formed (TwoDimension.py:159), compose / inverse follow
:475-477 and :494-498, exp / log maps :337-354 and :405-418."""
import numpy as np

TWO_PI = 2.0 * np.pi


def wrap(theta):
    """theta_to_pipi (src/utils/Functions.py:20-21)."""
    return (theta + np.pi) % TWO_PI - np.pi


def _rotate(th, xy):
    c, s = np.cos(th), np.sin(th)
    return np.stack([c * xy[..., 0] - s * xy[..., 1], s * xy[..., 0] + c * xy[..., 1]], axis=-1)


def se2_compose(a, b):
    """a * b for pose arrays (n, 3) / (3,)."""
    a, b = np.atleast_2d(a), np.atleast_2d(b)
    ath = wrap(a[:, 2])
    t = a[:, :2] + _rotate(ath, b[:, :2])
    return np.column_stack([t, wrap(ath + wrap(b[:, 2]))])


def se2_inverse(a):
    a = np.atleast_2d(a)
    ith = wrap(-wrap(a[:, 2]))
    return np.column_stack([-_rotate(ith, a[:, :2]), ith])


def se2_exp(v):
    """Exponential map of tangent vectors (n, 3) -> poses."""
    v = np.atleast_2d(v)
    w = v[:, 2]
    small = np.abs(w) < 1e-10
    ws = np.where(small, 1.0, w)
    ortho = _rotate(wrap(np.pi / 2), v[:, :2])
    t = (ortho - _rotate(wrap(ws), ortho)) / ws[:, None]
    t = np.where(small[:, None], v[:, :2], t)
    return np.column_stack([t, wrap(w)])


def se2_log(a):
    a = np.atleast_2d(a)
    w = wrap(a[:, 2])
    small = np.abs(w) < 1e-10
    ws = np.where(small, 1.0, w)
    c1, s = np.cos(ws) - 1.0, np.sin(ws)
    p = _rotate(wrap(np.pi / 2), _rotate(wrap(-ws), a[:, :2]) - a[:, :2])
    v = (ws / (c1 * c1 + s * s))[:, None] * p
    v = np.where(small[:, None], a[:, :2], v)
    return np.column_stack([v, w])


class SE2Pose:
    """Minimal single-pose value type with the reference's constructor / accessors (TwoDimension.py:303-330)."""

    dim = 3

    def __init__(self, x=None, y=None, theta=None):
        self.x = 0.0 if x is None else float(x)
        self.y = 0.0 if y is None else float(y)
        self.theta = float(wrap(0.0 if theta is None else float(theta)))

    @classmethod
    def by_array(cls, arr):
        return cls(arr[0], arr[1], arr[2])

    @classmethod
    def by_exp_map(cls, vector):
        return cls.by_array(se2_exp(np.asarray(vector, float))[0])

    @property
    def array(self):
        return np.array([self.x, self.y, self.theta])

    def inverse(self):
        return SE2Pose.by_array(se2_inverse(self.array)[0])

    def log_map(self):
        return se2_log(self.array)[0]

    def __mul__(self, other):
        return SE2Pose.by_array(se2_compose(self.array, other.array)[0])

    def __truediv__(self, other):
        return self * other.inverse()

    def __str__(self):
        return f"Pose2{{x: {self.x}, y: {self.y}, theta: {self.theta}}}"
