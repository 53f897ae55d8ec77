"""forces_zoo: every force / corrective / action-space kind of the hot path that no shipped
BASELINE scene reaches, in one small scene, so that the golden trajectory recorded from the
reference pins them:

  * pairwise `Gravity` (gravity.py:25-60), symmetric between two layers and asymmetric within
    one layer -- the (i, i) pairs of `itertools.product` take its `dist == 0` branch;
  * `KineticFriction` (friction.py:11-33) incl. a sprite at rest (`velocity_norm == 0`), one of
    infinite mass (skipped by abstract_force.py:64-74) and the sign flips of a sprite that the
    constant deceleration overshoots;
  * `DistanceForce(spring_force_fn)` (distance_fn_force.py:15-89), asymmetric and symmetric;
  * `TetherZippedLayers` (tether_physics.py:143-201) with and without `update_angle_vel`;
  * the `SetPosition` action space with inertia (set_position.py:13-47).
"""

import collections

import numpy as np

from moog import action_spaces
from moog import observers
from moog import physics as physics_lib
from moog import sprite
from moog import tasks


def get_config(level=None):
    del level

    def state_initializer():
        rng = np.random
        hsv = lambda k: dict(c0=0.09 * k, c1=1., c2=1.)
        planets = [sprite.Sprite(x=0.3, y=0.6, shape='circle', scale=0.08, mass=2., **hsv(0)),
                   sprite.Sprite(x=0.7, y=0.35, shape='circle', scale=0.1, mass=3., **hsv(1))]
        moons = [sprite.Sprite(x=0.3, y=0.6, shape='triangle', scale=0.04, x_vel=0.01, **hsv(2)),   # on planet 0
                 sprite.Sprite(x=0.5 + 0.05 * rng.rand(), y=0.8, shape='square', scale=0.04, y_vel=-0.01,
                               angle=0.4, angle_vel=0.1, **hsv(3)),
                 sprite.Sprite(x=0.55, y=0.2 + 0.05 * rng.rand(), shape='star_5', scale=0.05, x_vel=-0.02, **hsv(4))]
        sliders = [sprite.Sprite(x=0.15, y=0.15, shape='pentagon', scale=0.05, x_vel=0.013, y_vel=0.004, **hsv(5)),
                   sprite.Sprite(x=0.85, y=0.85, shape='hexagon', scale=0.05, **hsv(6)),                   # at rest
                   sprite.Sprite(x=0.85, y=0.15, shape='square', scale=0.05, x_vel=-0.01, mass=np.inf, **hsv(7)),
                   sprite.Sprite(x=0.15, y=0.85, shape='spoke_4', scale=0.05, x_vel=0.0007 * (1 + rng.rand()),
                                 y_vel=-0.0004, **hsv(8))]
        anchors = [sprite.Sprite(x=0.5, y=0.5, shape='square', scale=0.03, mass=np.inf, **hsv(9))]
        bobs = [sprite.Sprite(x=0.5, y=0.75, shape='circle', scale=0.04, x_vel=0.015, **hsv(10)),
                sprite.Sprite(x=0.4 - 0.05 * rng.rand(), y=0.45, shape='triangle', scale=0.04, mass=0.5, **hsv(1))]
        left = [sprite.Sprite(x=0.2, y=0.4, shape='square', scale=0.05, x_vel=0.01, y_vel=0.005, angle_vel=0.05, **hsv(2)),
                sprite.Sprite(x=0.8, y=0.6, shape='triangle', scale=0.05, y_vel=-0.01, mass=2., **hsv(3))]
        right = [sprite.Sprite(x=0.27, y=0.43, shape='circle', scale=0.03, **hsv(4)),
                 sprite.Sprite(x=0.74, y=0.66, shape='star_4', scale=0.04, x_vel=0.01 * rng.rand(), angle=1., **hsv(5))]
        carriers = [sprite.Sprite(x=0.6, y=0.9, shape='hexagon', scale=0.04, x_vel=-0.012, **hsv(6))]
        cargo = [sprite.Sprite(x=0.6, y=0.93, shape='square', scale=0.03, y_vel=-0.006, mass=0.25, **hsv(7))]
        cursor = [sprite.Sprite(x=0.5, y=0.1, shape='spoke_6', scale=0.05, **hsv(8))]
        return collections.OrderedDict([
            ('planets', planets), ('moons', moons), ('sliders', sliders), ('anchors', anchors), ('bobs', bobs),
            ('left', left), ('right', right), ('carriers', carriers), ('cargo', cargo), ('cursor', cursor)])

    physics = physics_lib.Physics(
        (physics_lib.Gravity(g=-0.01, symmetric=True), 'planets', 'moons'),
        (physics_lib.Gravity(g=-0.004, symmetric=False), 'moons', 'moons'),
        (physics_lib.KineticFriction(coeff_friction=0.0011), 'sliders'),
        (physics_lib.DistanceForce(physics_lib.spring_force_fn(0.05, equilibrium=0.2), symmetric=False), 'anchors', 'bobs'),
        (physics_lib.DistanceForce(physics_lib.spring_force_fn(0.02, equilibrium=0.1), symmetric=True), 'bobs', 'bobs'),
        corrective_physics=[
            physics_lib.TetherZippedLayers(('left', 'right'), update_angle_vel=True),
            physics_lib.TetherZippedLayers(('carriers', 'cargo'), update_angle_vel=False),
        ],
        updates_per_env_step=4)

    return {
        'state_initializer': state_initializer,
        'physics': physics,
        'task': tasks.CompositeTask(timeout_steps=60),
        'action_space': action_spaces.SetPosition(action_layers='cursor', inertia=0.3),
        'observers': {'image': observers.PILRenderer(image_size=(64, 64), anti_aliasing=1, color_to_rgb='hsv_to_rgb')},
        'game_rules': (),
    }
