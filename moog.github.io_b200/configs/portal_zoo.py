"""portal_zoo: the two layer-level rules of the reference that move sprites around the state --
`Portal` (game_rules/portal.py:14-76: a sprite whose position enters a portal sprite reappears at
the partner portal and cannot teleport again until it has left every portal; one pair of square
portals, one pair of circular ones, which take the disc test of sprite.py:432-440) and `ChangeLayer`
(change_layer.py:11-45: sprites that cross x = 0.5 move from the frictionless layer `left` to the
layer `right`, which has Drag; the receiving layer starts empty).
"""

import collections

import numpy as np

from moog import action_spaces
from moog import game_rules
from moog import observers
from moog import physics as physics_lib
from moog import sprite
from moog import tasks


def get_config(level=None):
    del level

    def state_initializer():
        rng = np.random
        portals = [
            sprite.Sprite(x=0.2, y=0.8, shape='square', scale=0.12, c0=0.75, c1=1., c2=1.),
            sprite.Sprite(x=0.8, y=0.2, shape='square', scale=0.12, angle=0.4, c0=0.75, c1=1., c2=0.6),
            sprite.Sprite(x=0.2, y=0.2, shape='circle', scale=0.12, c0=0.55, c1=1., c2=1.),
            sprite.Sprite(x=0.8, y=0.8, shape='circle', scale=0.12, c0=0.55, c1=1., c2=0.6),
        ]
        movers = [
            sprite.Sprite(x=0.5, y=0.8, shape='triangle', scale=0.05, x_vel=-0.03, c0=0.1, c1=1., c2=1.),
            sprite.Sprite(x=0.2 + 0.01 * rng.rand(), y=0.5, shape='star_5', scale=0.05, y_vel=-0.025, c0=0.2, c1=1., c2=1.),
            sprite.Sprite(x=0.6, y=0.6, shape='circle', scale=0.04, x_vel=0.021, y_vel=0.02, c0=0.3, c1=1., c2=1.),
        ]
        left = [
            sprite.Sprite(x=0.1 + 0.12 * k, y=0.35 + 0.1 * k + 0.01 * rng.rand(), shape='pentagon', scale=0.05,
                          x_vel=0.02 + 0.01 * k, angle_vel=0.05, c0=0.9, c1=0.5 + 0.1 * k, c2=1.)
            for k in range(4)]
        return collections.OrderedDict([
            ('portals', portals), ('movers', movers), ('left', left), ('right', []), ('agent', [])])

    physics = physics_lib.Physics(
        (physics_lib.Drag(coeff_friction=0.3), 'right'),
        updates_per_env_step=2)

    rules = (
        game_rules.Portal(teleporting_layer='movers', portal_layer='portals'),
        game_rules.ChangeLayer('left', 'right', filter_fn=lambda s: s.x > 0.5),
    )

    return {
        'state_initializer': state_initializer,
        'physics': physics,
        'task': tasks.CompositeTask(timeout_steps=60),
        'action_space': action_spaces.Grid(action_layers='agent'),
        'observers': {'image': observers.PILRenderer(image_size=(64, 64), anti_aliasing=1, color_to_rgb='hsv_to_rgb')},
        'game_rules': rules,
    }
