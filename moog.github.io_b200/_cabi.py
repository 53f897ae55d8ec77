"""Host-side geometry entry points of libmoog_b200.so for the host `Sprite`
(moog/sprite.py:442-484): only used while the state initializer builds
episodes (rejection sampling); the step itself runs in the CUDA kernels."""
import ctypes
import math

import numpy as np

from . import capi


def _closed(sprite):
    return np.ascontiguousarray(sprite.path.vertices, dtype=np.float64)


def host_sprites_overlap(a, b):
    """Sprite.overlaps_sprite (sprite.py:462-484)."""
    d = a.position - b.position
    if math.sqrt(d[0] * d[0] + d[1] * d[1]) > a.max_radius + b.max_radius:
        return False
    pa, pb = _closed(a), _closed(b)
    return bool(capi.lib().moog_host_paths_overlap(
        pa.ctypes.data_as(ctypes.c_void_p), len(pa),
        pb.ctypes.data_as(ctypes.c_void_p), len(pb)))


def host_sprite_contains_points(sprite, points):
    """Sprite.contains_points (sprite.py:442-460)."""
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
    if sprite.is_symmetric_circle:
        d = pts - sprite.position
        return np.sqrt(np.sum(d * d, axis=1)) <= sprite.max_radius
    path = _closed(sprite)
    out = np.zeros(len(pts), dtype=np.uint8)
    capi.lib().moog_host_points_in_path(
        pts.ctypes.data_as(ctypes.c_void_p), len(pts),
        path.ctypes.data_as(ctypes.c_void_p), len(path),
        out.ctypes.data_as(ctypes.c_void_p))
    return out.astype(bool)
