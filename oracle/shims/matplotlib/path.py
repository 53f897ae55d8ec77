"""Stand-in for `matplotlib.path.Path` restricted to what MOOG calls.

TEST INFRASTRUCTURE ONLY.  matplotlib (pinned 3.10.0 in the reference's
setup.py:39-46) is not installed in this image, so the three `_path` C++
routines the reference reaches through `Path` are restated here from the
published algorithm in matplotlib's src/_path.h:

  * point_in_path_impl       -> Path.contains_point(s)   (crossing number)
  * segments_intersect / path_intersects_path / path_in_path
                             -> Path.intersects_path(a, b, filled=True)

Call sites in the reference: moog/sprite.py:396,418,438,458,482-483 and
moog/physics/collisions.py:146.  Pinned by the reference's own KATs
(tests/moog/physics/test_collisions.py, test_tether_physics.py) which pass
with this shim; matplotlib itself is not available to diff against, so the
matplotlib-internal tie-break conventions are "recalled, KAT-confirmed".

When oracle/_build/libmoog_oracle.so exists the two hot predicates are routed
to its C restatement (same algorithm, ~100x faster) so that reference timings
taken through this shim are not handicapped by pure-Python geometry.
"""

import ctypes
import os

import numpy as np

_RTOL = 1e-10
_ATOL = 1e-13

_LIB = None
_lib_path = os.path.join(
    os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
    '_build', 'libmoog_oracle.so')
if os.path.exists(_lib_path) and not os.environ.get('MOOG_SHIM_PURE_PYTHON'):
    try:
        _LIB = ctypes.CDLL(_lib_path)
        _dp = ctypes.POINTER(ctypes.c_double)
        _LIB.orc_path_intersects_filled.argtypes = [
            _dp, ctypes.c_int, _dp, ctypes.c_int]
        _LIB.orc_path_intersects_filled.restype = ctypes.c_int
        _LIB.orc_points_in_path.argtypes = [
            _dp, ctypes.c_int, _dp, ctypes.c_int,
            ctypes.POINTER(ctypes.c_uint8)]
        _LIB.orc_points_in_path.restype = None
    except (OSError, AttributeError):
        _LIB = None


def _isclose(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) <= np.maximum(
        _RTOL * np.maximum(np.abs(a), np.abs(b)), _ATOL)


def _effective_segments(v):
    """Segments of an open polyline after the zero-length-skip rule.

    A segment whose squared length isclose to 0 is skipped and its START
    vertex is kept as the start of the next segment.
    """
    starts, ends = [], []
    x1, y1 = v[0]
    for k in range(1, len(v)):
        x2, y2 = v[k]
        d2 = (x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2)
        if abs(d2) <= max(_RTOL * abs(d2), _ATOL):
            continue
        starts.append((x1, y1))
        ends.append((x2, y2))
        x1, y1 = x2, y2
    return (np.array(starts, dtype=np.float64).reshape(-1, 2),
            np.array(ends, dtype=np.float64).reshape(-1, 2))


def _any_segments_intersect(a, b):
    s1, e1 = _effective_segments(a)
    s2, e2 = _effective_segments(b)
    if len(s1) == 0 or len(s2) == 0:
        return False
    x1 = s1[:, None, 0]
    y1 = s1[:, None, 1]
    x2 = e1[:, None, 0]
    y2 = e1[:, None, 1]
    x3 = s2[None, :, 0]
    y3 = s2[None, :, 1]
    x4 = e2[None, :, 0]
    y4 = e2[None, :, 1]

    den = ((y4 - y3) * (x2 - x1)) - ((x4 - x3) * (y2 - y1))
    den_zero = _isclose(den, 0.0)

    # Degenerate (parallel) branch.
    t_area = (x2 * y3 - x3 * y2) - x1 * (y3 - y2) + y1 * (x3 - x2)
    collinear = _isclose(t_area, 0.0)
    vertical = (x1 == x2) & (x2 == x3)

    def _ovl(p1, p2, p3, p4):
        lo12 = np.minimum(p1, p2)
        hi12 = np.maximum(p1, p2)
        lo34 = np.minimum(p3, p4)
        hi34 = np.maximum(p3, p4)
        return (((lo12 <= lo34) & (lo34 <= hi12)) |
                ((lo34 <= lo12) & (lo12 <= hi34)))

    par_hit = collinear & np.where(
        vertical, _ovl(y1, y2, y3, y4), _ovl(x1, x2, x3, x4))

    n1 = ((x4 - x3) * (y1 - y3)) - ((y4 - y3) * (x1 - x3))
    n2 = ((x2 - x1) * (y1 - y3)) - ((y2 - y1) * (x1 - x3))
    with np.errstate(divide='ignore', invalid='ignore'):
        u1 = n1 / den
        u2 = n2 / den
        reg_hit = (((u1 > 0.0) | _isclose(u1, 0.0)) &
                   ((u1 < 1.0) | _isclose(u1, 1.0)) &
                   ((u2 > 0.0) | _isclose(u2, 0.0)) &
                   ((u2 < 1.0) | _isclose(u2, 1.0)))
    return bool(np.any(np.where(den_zero, par_hit, reg_hit)))


def _points_in_path(points, v):
    """Crossing-number test of `points` [n,2] against closed polygon `v` [m,2]."""
    n = len(points)
    if len(v) < 3:
        return np.zeros(n, dtype=bool)
    x0 = v[:, 0][None, :]
    y0 = v[:, 1][None, :]
    x1 = np.roll(v[:, 0], -1)[None, :]
    y1 = np.roll(v[:, 1], -1)[None, :]
    tx = points[:, 0][:, None]
    ty = points[:, 1][:, None]
    f0 = y0 >= ty
    f1 = y1 >= ty
    cond = ((y1 - ty) * (x0 - x1) >= (x1 - tx) * (y0 - y1)) == f1
    toggles = (f0 != f1) & cond
    inside = (np.sum(toggles, axis=1) % 2).astype(bool)
    finite = np.isfinite(points[:, 0]) & np.isfinite(points[:, 1])
    return inside & finite


def _as_c(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class Path(object):
    """Polyline with `vertices` [n, 2] float64 and no codes."""

    def __init__(self, vertices, codes=None):
        self.vertices = np.array(vertices, dtype=np.float64).reshape(-1, 2)
        self.codes = None

    def __len__(self):
        return len(self.vertices)

    def contains_points(self, points, transform=None, radius=0.0):
        points = np.asarray(points, dtype=np.float64).reshape(-1, 2)
        if _LIB is not None:
            p, pp = _as_c(points)
            v, vp = _as_c(self.vertices)
            out = np.zeros(len(p), dtype=np.uint8)
            _LIB.orc_points_in_path(
                pp, len(p), vp, len(v),
                out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
            return out.astype(bool)
        return _points_in_path(points, self.vertices)

    def contains_point(self, point, transform=None, radius=0.0):
        return bool(self.contains_points(np.asarray(point).reshape(1, 2))[0])

    def contains_path(self, other, transform=None):
        """True iff every vertex of `other` lies inside self (path_in_path)."""
        if len(other.vertices) < 3:
            return False
        return bool(np.all(self.contains_points(other.vertices)))

    def intersects_path(self, other, filled=True):
        if _LIB is not None and filled:
            a, ap = _as_c(self.vertices)
            b, bp = _as_c(other.vertices)
            return bool(
                _LIB.orc_path_intersects_filled(ap, len(a), bp, len(b)))
        a = self.vertices
        b = other.vertices
        hit = False
        if len(a) >= 2 and len(b) >= 2:
            hit = _any_segments_intersect(a, b)
        if filled:
            # path_in_path(self, other): self inside other; then the converse.
            if not hit:
                hit = other.contains_path(self)
            if not hit:
                hit = self.contains_path(other)
        return hit
