"""Unit-area polygon generators (reference: moog/polygons.py:10-92).

Each generator returns an [n, 2] float64 vertex array, counter-clockwise,
scaled so that the enclosed area is 1.
"""

import numpy as np


def _ring(radius, angles):
    """Points at `radius` (scalar or [n]) and `angles` [n] -> [n, 2]."""
    angles = np.asarray(angles, dtype=np.float64)
    radius = np.broadcast_to(np.asarray(radius, dtype=np.float64), angles.shape)
    return np.stack([radius * np.cos(angles), radius * np.sin(angles)], axis=1)


def polygon(num_sides, theta_0=0.):
    """Regular `num_sides`-gon with first vertex at angle `theta_0`."""
    step = 2 * np.pi / num_sides
    verts = _ring(1, np.arange(num_sides) * step + theta_0)
    area = num_sides * np.sin(step / 2) * np.cos(step / 2)
    return verts / np.sqrt(area)


def star(num_sides, point_height=1, theta_0=0.):
    """Star with `num_sides` points of height `point_height` above the unit
    inscribed circle; vertices alternate inner / outer."""
    step = 2 * np.pi / num_sides
    outer_r = 1 + point_height
    k = np.arange(num_sides)
    verts = np.empty((2 * num_sides, 2))
    verts[0::2] = _ring(1, k * step + theta_0)
    verts[1::2] = _ring(outer_r, (k + 0.5) * step + theta_0)
    area = outer_r * num_sides * np.sin(step / 2)
    return verts / np.sqrt(area)


def spokes(num_sides, spoke_height=1, theta_0=0.):
    """Like `star` but with rectangular spokes: 3 vertices per side."""
    step = 2 * np.pi / num_sides
    k = np.arange(num_sides)
    hub = _ring(1, k * step + theta_0)
    before = _ring(spoke_height, (k - 0.5) * step + theta_0)
    after = _ring(spoke_height, (k + 0.5) * step + theta_0)
    verts = np.empty((3 * num_sides, 2))
    verts[0::3] = before + hub
    verts[1::3] = hub
    verts[2::3] = after + hub
    area = num_sides * np.sin(step / 2) * (2 + np.cos(step / 2))
    return verts / np.sqrt(area)
