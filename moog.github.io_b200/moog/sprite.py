"""Host-side Sprite: the unit of state a MOOG state initializer produces.

Reference: moog/sprite.py:229-675.  A Sprite here is a factor record plus the
derived quantities the device needs (COM-centred base outline, rotational
inertia per unit mass, circumscribed radius).  World vertices are computed on
demand from the factors; nothing is cached incrementally (the device does the
same, see DESIGN.md "geometry").
"""

import collections
import math

import numpy as np

from moog import shapes

GLOBAL_SPRITE_COUNT = 0


class _Outline(object):
    """Minimal stand-in for the `Path` object the reference keeps in
    `Sprite._shape_path` / `Sprite._path` (only `.vertices` is used)."""

    __slots__ = ('vertices',)

    def __init__(self, vertices):
        self.vertices = vertices


def _polygon_moments(verts):
    """Signed area, centroid and (I_x, I_y) about the origin of a polygon.

    Fan of triangles (origin, v_i, v_{i+1}); sums are accumulated in vertex
    order (sprite.py:360-379).
    """
    nxt = np.roll(verts, -1, axis=0)
    cross = verts[:, 0] * nxt[:, 1] - verts[:, 1] * nxt[:, 0]
    tri_inertia = ((1. / 12.) * cross)[:, None] * (
        verts * verts + nxt * nxt + verts * nxt)
    tri_area = cross / 2.
    tri_centroid = (verts + nxt) / 3.
    # cumsum accumulates left to right, like the reference's Python loop.
    inertia = np.cumsum(tri_inertia, axis=0)[-1]
    area = np.cumsum(tri_area)[-1]
    centroid = np.cumsum(tri_centroid * tri_area[:, None], axis=0)[-1] / area
    return area, centroid, inertia


class Sprite(object):
    """Polygon sprite parameterised by the 15 MOOG factors."""

    FACTOR_NAMES = (
        'x', 'y', 'shape', 'angle', 'scale', 'aspect_ratio', 'c0', 'c1', 'c2',
        'opacity', 'x_vel', 'y_vel', 'angle_vel', 'mass', 'metadata',
    )

    _CUSTOM_SHAPE = 'custom'
    _CIRCLE_NAME = 'circle'

    def __init__(self, x=0.5, y=0.5, shape='square', angle=0., scale=1.,
                 aspect_ratio=1., c0=0, c1=0, c2=0, opacity=255, x_vel=0.,
                 y_vel=0., angle_vel=0., mass=1., metadata=None):
        global GLOBAL_SPRITE_COUNT
        self._position = np.array([x, y])
        self._angle = float(angle)
        self._scale = float(scale)
        self._aspect_ratio = float(aspect_ratio)
        self._color = (c0, c1, c2)
        self._opacity = opacity
        self._velocity = np.array([x_vel, y_vel])
        self._angle_vel = angle_vel
        self._mass = mass
        self.metadata = metadata
        self.shape = shape
        self._id = GLOBAL_SPRITE_COUNT
        GLOBAL_SPRITE_COUNT += 1

    # -- geometry ---------------------------------------------------------
    def _install_outline(self, outline):
        """Centre `outline` on its centroid, fix its winding, derive inertia.

        sprite.py:329-409: clockwise outlines are reversed; the sprite's
        position is then shifted by the RAW centroid (not scaled/rotated).
        """
        outline = np.array(outline, dtype=np.float64)
        area, centroid, inertia = _polygon_moments(outline)
        if area < 0:
            outline = outline[::-1]
            inertia = inertia * -1.
            area = area * -1.
        closed = np.concatenate((outline, outline[:1]), axis=0)
        self._shape_path = _Outline(closed - centroid)
        inertia = inertia - area * np.square(centroid)
        self._x_y_rotational_inertia = inertia / area
        self._refresh_derived()
        self._position = self._position + centroid
        self._just_set_shape = True

    def _matrix(self):
        sx = self._scale
        sy = self._scale * self._aspect_ratio
        c, s = math.cos(self._angle), math.sin(self._angle)
        return c * sx, -(s * sy), s * sx, c * sy

    def _world(self, closed):
        m00, m01, m10, m11 = self._matrix()
        base = self._shape_path.vertices if closed else (
            self._shape_path.vertices[:-1])
        out = np.empty_like(base)
        out[:, 0] = m00 * base[:, 0] + m01 * base[:, 1] + self._position[0]
        out[:, 1] = m10 * base[:, 0] + m11 * base[:, 1] + self._position[1]
        return out

    def _refresh_derived(self):
        """sprite.py:411-424: circumscribed radius; inertia scales by
        (scale, scale*aspect)^2 -- compounding on every call, as upstream."""
        rel = self._world(closed=False) - self._position
        self._max_radius = np.max(np.sqrt(np.sum(rel * rel, axis=1)))
        xy_scale = np.array([self._scale, self._scale * self._aspect_ratio])
        self._x_y_rotational_inertia = (
            self._x_y_rotational_inertia * np.square(xy_scale))

    @property
    def vertices(self):
        return self._world(closed=False)

    @property
    def path(self):
        return _Outline(self._world(closed=True))

    @property
    def max_radius(self):
        return self._max_radius

    @property
    def is_symmetric_circle(self):
        return self.shape == Sprite._CIRCLE_NAME and self.aspect_ratio == 1

    def overlaps_sprite(self, other):
        """Filled-polygon overlap, evaluated by the native library's host
        entry point (same predicate the device kernels use)."""
        from moog_b200 import _cabi
        return _cabi.host_sprites_overlap(self, other)

    def contains_points(self, points):
        from moog_b200 import _cabi
        return _cabi.host_sprite_contains_points(self, np.asarray(points))

    def contains_point(self, point):
        if self.is_symmetric_circle:
            d = np.asarray(point, dtype=float) - self._position
            return bool(math.sqrt(d[0] * d[0] + d[1] * d[1]) < self._max_radius)
        return bool(self.contains_points(np.asarray(point).reshape(1, 2))[0])

    def update_pos_from_vel(self, delta_t):
        self.position = self._position + delta_t * self._velocity
        if self._angle_vel:
            self.angle = self._angle + delta_t * self._angle_vel

    # -- factors ----------------------------------------------------------
    @property
    def shape(self):
        return self._shape

    @shape.setter
    def shape(self, shape):
        if isinstance(shape, str) and shape in shapes.SHAPES:
            self._shape = shape
            self._install_outline(shapes.SHAPES[shape])
        else:
            self._shape = Sprite._CUSTOM_SHAPE
            self._install_outline(shape)

    @property
    def x(self):
        return self._position[0]

    @property
    def y(self):
        return self._position[1]

    @property
    def position(self):
        return self._position

    @position.setter
    def position(self, pos):
        if pos is self._position:
            raise ValueError(
                'Cannot call in-place operations on sprite.position.')
        self._position = pos if isinstance(pos, np.ndarray) else np.array(pos)

    @property
    def velocity(self):
        return self._velocity

    @velocity.setter
    def velocity(self, vel):
        self._velocity = vel if isinstance(vel, np.ndarray) else np.array(vel)

    @property
    def x_vel(self):
        return self._velocity[0]

    @property
    def y_vel(self):
        return self._velocity[1]

    @property
    def angle(self):
        return self._angle

    @angle.setter
    def angle(self, a):
        self._angle = a

    @property
    def angle_vel(self):
        return self._angle_vel

    @angle_vel.setter
    def angle_vel(self, w):
        self._angle_vel = w

    @property
    def scale(self):
        return self._scale

    @scale.setter
    def scale(self, s):
        self._scale = s
        self._refresh_derived()

    @property
    def aspect_ratio(self):
        return self._aspect_ratio

    @aspect_ratio.setter
    def aspect_ratio(self, a):
        self._aspect_ratio = a
        self._refresh_derived()

    @property
    def mass(self):
        return self._mass

    @mass.setter
    def mass(self, m):
        self._mass = m

    @property
    def moment_of_inertia(self):
        return sum(self._mass * self._x_y_rotational_inertia)

    @property
    def color(self):
        return self._color

    def _set_color(self, index, value):
        c = list(self._color)
        c[index] = value
        self._color = tuple(c)

    c0 = property(lambda self: self._color[0],
                  lambda self, v: self._set_color(0, v))
    c1 = property(lambda self: self._color[1],
                  lambda self, v: self._set_color(1, v))
    c2 = property(lambda self: self._color[2],
                  lambda self, v: self._set_color(2, v))

    @property
    def opacity(self):
        return self._opacity

    @opacity.setter
    def opacity(self, o):
        self._opacity = o

    @property
    def just_set_shape(self):
        return self._just_set_shape

    @just_set_shape.setter
    def just_set_shape(self, flag):
        self._just_set_shape = flag

    @property
    def id(self):
        return self._id

    @property
    def factors(self):
        return collections.OrderedDict(
            (name, getattr(self, name)) for name in Sprite.FACTOR_NAMES)
