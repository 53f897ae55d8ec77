"""Stand-in for `matplotlib.transforms.Affine2D` restricted to MOOG's usage.

TEST INFRASTRUCTURE ONLY (see path.py).  Restates matplotlib 3.10
lib/matplotlib/transforms.py `Affine2D`: a 3x3 float64 matrix; `scale`,
`rotate`, `translate` left-multiply the current matrix; `rotate_around` is
translate(-x,-y).rotate(theta).translate(x,y); `a + b` applies a then b.
Call sites: moog/sprite.py:390,414-418,537-539,631-632 and
moog/physics/collisions.py:83-95.
"""

import math

import numpy as np

from .path import Path


class Affine2D(object):
    def __init__(self, matrix=None):
        if matrix is None:
            self._mtx = np.identity(3)
        else:
            self._mtx = np.array(matrix, dtype=np.float64)

    def get_matrix(self):
        return self._mtx

    def scale(self, sx, sy=None):
        if sy is None:
            sy = sx
        m = self._mtx
        m[0, 0] *= sx
        m[0, 1] *= sx
        m[0, 2] *= sx
        m[1, 0] *= sy
        m[1, 1] *= sy
        m[1, 2] *= sy
        return self

    def rotate(self, theta):
        a = math.cos(theta)
        b = math.sin(theta)
        m = self._mtx
        (xx, xy, x0), (yx, yy, y0), _ = m.tolist()
        m[0, 0] = a * xx - b * yx
        m[0, 1] = a * xy - b * yy
        m[0, 2] = a * x0 - b * y0
        m[1, 0] = b * xx + a * yx
        m[1, 1] = b * xy + a * yy
        m[1, 2] = b * x0 + a * y0
        return self

    def translate(self, tx, ty):
        self._mtx[0, 2] += tx
        self._mtx[1, 2] += ty
        return self

    def rotate_around(self, x, y, theta):
        return self.translate(-x, -y).rotate(theta).translate(x, y)

    def __add__(self, other):
        # "self, then other"
        return Affine2D(np.dot(other._mtx, self._mtx))

    def transform(self, points):
        points = np.asarray(points, dtype=np.float64)
        m = self._mtx
        x = points[..., 0]
        y = points[..., 1]
        out = np.empty_like(points)
        out[..., 0] = m[0, 0] * x + m[0, 1] * y + m[0, 2]
        out[..., 1] = m[1, 0] * x + m[1, 1] * y + m[1, 2]
        return out

    def transform_path(self, path):
        return Path(self.transform(path.vertices))
