"""synthetic-32: BASELINE.json config 5 (SURVEY section 8d scene 5).

4 border walls + 28 polygon sprites cycling through 7 shapes on a jittered
6x5 grid (disjoint by construction), random angle / velocity / angular
velocity, symmetric rotational Collision between sprites and an asymmetric one
against the walls, K=10, no RNG inside step().
"""

import collections

import numpy as np

from moog import action_spaces
from moog import observers
from moog import physics as physics_lib
from moog import shapes
from moog import sprite as sprite_lib
from moog import tasks

_SHAPES = ('triangle', 'square', 'pentagon', 'hexagon', 'star_5', 'spoke_4',
           'circle')


def get_config(level=None):
    level = level or {}
    image_size = level.get('image_size', (64, 64))
    render = level.get('render', True)
    timeout = level.get('timeout_steps', 200)

    def state_initializer():
        walls = shapes.border_walls(visible_thickness=0.05, c0=0., c1=0.,
                                    c2=0.5)
        sprites = []
        k = 0
        for gy in range(5):
            for gx in range(6):
                if k >= 28:
                    break
                cx = 0.12 + (gx + 0.5) * (0.76 / 6) + np.random.uniform(
                    -0.012, 0.012)
                cy = 0.12 + (gy + 0.5) * (0.76 / 5) + np.random.uniform(
                    -0.012, 0.012)
                sprites.append(sprite_lib.Sprite(
                    x=cx, y=cy, shape=_SHAPES[k % len(_SHAPES)],
                    scale=np.random.uniform(0.05, 0.08),
                    angle=np.random.uniform(0., 2 * np.pi),
                    x_vel=np.random.uniform(-0.02, 0.02),
                    y_vel=np.random.uniform(-0.02, 0.02),
                    angle_vel=np.random.uniform(-0.05, 0.05),
                    c0=(k % 7) / 7., c1=1., c2=1., mass=1.))
                k += 1
        return collections.OrderedDict([
            ('walls', walls),
            ('sprites', sprites),
            ('agent', []),
        ])

    physics = physics_lib.Physics(
        (physics_lib.Collision(elasticity=1., symmetric=True,
                               update_angle_vel=True), 'sprites', 'sprites'),
        (physics_lib.Collision(elasticity=1., symmetric=False,
                               update_angle_vel=True), 'sprites', 'walls'),
        updates_per_env_step=10,
    )
    task = tasks.CompositeTask(timeout_steps=timeout)
    action_space = action_spaces.Grid(action_layers='agent')
    obs = {}
    if render:
        obs['image'] = observers.PILRenderer(
            image_size=image_size, anti_aliasing=1, color_to_rgb='hsv_to_rgb')
    return {
        'state_initializer': state_initializer,
        'physics': physics,
        'task': task,
        'action_space': action_space,
        'observers': obs,
        'game_rules': (),
    }
