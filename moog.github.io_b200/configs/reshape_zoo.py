"""reshape_zoo: rules that re-parameterise sprites: assignments of
`scale`, `aspect_ratio` and `angle` by modifiers (sprite.py:516-558), each of which makes the
reference rebuild the outline from the shape (`_set_path`, :411-424) and multiply the rotational
inertia by the squared scales AGAIN -- so the inertia that the rotational Collision uses afterwards
compounds.  Polygons bounce around an arena (symmetric rotational Collision among themselves,
asymmetric against the walls); every contact with a wall shrinks the polygon and makes it more
oblong, a timed rule resets the aspect ratios and another one snaps the angles.
"""

import collections

import numpy as np

from moog import action_spaces
from moog import game_rules
from moog import observers
from moog import physics as physics_lib
from moog import shapes
from moog import sprite
from moog import tasks


def get_config(level=None):
    del level

    def state_initializer():
        rng = np.random
        walls = shapes.border_walls(visible_thickness=0.05, c0=0., c1=0., c2=0.5)
        kinds = ['triangle', 'square', 'pentagon', 'star_5', 'spoke_4', 'hexagon']
        blobs = [
            sprite.Sprite(x=0.2 + 0.3 * (k % 3) + 0.02 * rng.rand(), y=0.3 + 0.4 * (k // 3), shape=kinds[k],
                          scale=0.12, angle=0.5 * k, x_vel=0.04 * rng.rand() - 0.02, y_vel=0.04 * rng.rand() - 0.02,
                          angle_vel=0.1 * rng.rand() - 0.05, c0=0.15 * k, c1=1., c2=1.)
            for k in range(6)]
        blobs.append(sprite.Sprite(x=0.5, y=0.5, shape=np.array([[-1., -0.6], [1.2, -0.4], [0.3, 0.9]]), scale=0.08,
                                   x_vel=0.015, y_vel=-0.02, c0=0.9, c1=1., c2=1.))     # a custom outline
        return collections.OrderedDict([('walls', walls), ('blobs', blobs), ('agent', [])])

    physics = physics_lib.Physics(
        (physics_lib.Collision(elasticity=1., symmetric=True, update_angle_vel=True), 'blobs', 'blobs'),
        (physics_lib.Collision(elasticity=1., symmetric=False, update_angle_vel=True), 'blobs', 'walls'),
        updates_per_env_step=5)

    def _squash(s):
        s.scale = 0.93 * s.scale
        s.aspect_ratio = 1.1 * s.aspect_ratio

    def _round(s):
        s.aspect_ratio = 1.

    def _snap(s):
        s.angle = 0.25

    rules = (
        game_rules.ModifyOnContact(layers_0='blobs', layers_1='walls', modifier_0=_squash),
        game_rules.TimedRule((12, 13), game_rules.ModifySprites('blobs', _round)),
        game_rules.TimedRule((20, 22), game_rules.ModifySprites('blobs', _snap, filter_fn=lambda s: s.scale < 0.11)),
    )

    return {
        'state_initializer': state_initializer,
        'physics': physics,
        'task': tasks.CompositeTask(timeout_steps=60),
        'action_space': action_spaces.Grid(action_layers='agent'),
        'observers': {'image': observers.PILRenderer(image_size=(64, 64), anti_aliasing=1, color_to_rgb='hsv_to_rgb')},
        'game_rules': rules,
    }
