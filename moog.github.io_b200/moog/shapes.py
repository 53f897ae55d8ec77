"""Named unit-area shapes and wall / grid helpers.

Reference: moog/shapes.py:11-24 (SHAPES), :27-77 (border_walls), :80-150
(grid_lines), :153-188 (circle_vertices, annulus_vertices).
"""

import numpy as np

from moog import polygons
from moog import sprite

_PI = np.pi

SHAPES = {
    'triangle': polygons.polygon(3, theta_0=_PI / 2),
    'square': polygons.polygon(4, theta_0=_PI / 4),
    'pentagon': polygons.polygon(5, theta_0=_PI / 2),
    'hexagon': polygons.polygon(6),
    'octagon': polygons.polygon(8),
    'circle': polygons.polygon(30),
    'star_4': polygons.star(4, theta_0=_PI / 4),
    'star_5': polygons.star(5, theta_0=_PI + _PI / 10),
    'star_6': polygons.star(6),
    'spoke_4': polygons.spokes(4, theta_0=_PI / 4),
    'spoke_5': polygons.spokes(5, theta_0=_PI + _PI / 10),
    'spoke_6': polygons.spokes(6),
}


def _rect(x_lo, x_hi, y_lo, y_hi):
    return np.array(
        [[x_lo, y_lo], [x_hi, y_lo], [x_hi, y_hi], [x_lo, y_hi]], dtype=float)


def border_walls(visible_thickness=0.05, total_thickness=0.5,
                 c0=0, c1=0, c2=0, opacity=255):
    """Four wall sprites framing [0, 1]^2: bottom, top, left, right.

    `visible_thickness` of each wall lies inside the frame, the rest of its
    `total_thickness` outside.
    """
    inner = visible_thickness
    outer = visible_thickness - total_thickness
    # Bottom wall, vertex order as in the reference (clockwise; Sprite flips it).
    bottom = np.array([[0., inner], [1., inner], [1., outer], [0., outer]])
    span = 1 + total_thickness - 2 * visible_thickness
    left = bottom[:, ::-1]
    outlines = [
        bottom,
        bottom + np.array([[0., span]]),
        left,
        left + np.array([[span, 0.]]),
    ]
    return [
        sprite.Sprite(x=0., y=0., shape=o, c0=c0, c1=c1, c2=c2, opacity=opacity)
        for o in outlines
    ]


def grid_lines(grid_x=0.4, grid_y=0.4, line_thickness=0.01, buffer_border=0.,
               c0=0, c1=0, c2=0, opacity=255):
    """Thin rectangles forming a background grid centred on (0.5, 0.5)."""
    half_t = 0.5 * line_thickness
    lo, hi = -1 * buffer_border, 1. + buffer_border
    n_x = int(np.floor((0.5 + buffer_border) / grid_x))
    n_y = int(np.floor((0.5 + buffer_border) / grid_y))
    xs = np.linspace(0.5 - n_x * grid_x, 0.5 + n_x * grid_x, 1 + 2 * n_x)
    ys = np.linspace(0.5 - n_y * grid_y, 0.5 + n_y * grid_y, 1 + 2 * n_y)
    outlines = [_rect(x - half_t, x + half_t, lo, hi) for x in xs]
    outlines += [_rect(lo, hi, y - half_t, y + half_t) for y in ys]
    return [
        sprite.Sprite(x=0., y=0., shape=o, c0=c0, c1=c1, c2=c2, opacity=opacity)
        for o in outlines
    ]


def circle_vertices(radius, num_sides=50):
    """`num_sides`-gon of the given radius about the origin, starting just
    past 12 o'clock and running clockwise."""
    thetas = np.linspace(2 * np.pi / num_sides, 2 * np.pi, num_sides)
    return radius * np.stack([np.sin(thetas), np.cos(thetas)], axis=1)


def annulus_vertices(inner_radius, outer_radius, num_sides=50):
    """Closed inner ring followed by the reversed closed outer ring."""
    def _closed(r):
        ring = circle_vertices(r, num_sides=num_sides)
        return np.concatenate((ring, ring[:1]), axis=0)
    return np.concatenate(
        (_closed(inner_radius), _closed(outer_radius)[::-1]), axis=0)
