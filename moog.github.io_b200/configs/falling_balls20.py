"""falling_balls-20: BASELINE.json config 2 (SURVEY section 8d scene 2).

The shipped moog_demos/example_configs/falling_balls.py with `num_sprites=20`
and `scale=0.05` (the shipped scale 0.1 cannot fit 20 disjoint balls in its
spawn box, falling_balls.py:44-52); everything else -- the 4 walls, the
asymmetric Collision(elasticity=.6, max_recursion_depth=2) on
('balls', ['balls', 'walls']), DownGravity(-0.001), K=20, timeout 100 and the
64x64 PILRenderer -- is unchanged.
"""

import collections

import numpy as np

from moog import action_spaces
from moog import observers
from moog import physics as physics_lib
from moog import sprite as sprite_lib
from moog import tasks
from moog.state_initialization import distributions as distribs
from moog.state_initialization import sprite_generators


def get_config(level=None):
    """level: None or dict(num_sprites=..., scale=..., image_size=...)."""
    level = level or {}
    num_sprites = level.get('num_sprites', 20)
    scale = level.get('scale', 0.05)
    image_size = level.get('image_size', (64, 64))

    ball_factors = distribs.Product(
        [distribs.Continuous('x', 0.25, 0.75),
         distribs.Continuous('y', 0.5, 0.9),
         distribs.Continuous('x_vel', -0.01, 0.01)],
        scale=scale, shape='circle', c0=0, c1=0, c2=255, mass=1.,
    )
    ball_generator = sprite_generators.generate_sprites(
        ball_factors, num_sprites=num_sprites)

    bottom_wall = [[-1, 0.1], [2, 0.1], [2, -1], [-1, -1]]
    left_wall = [[0.05, -0.1], [0.05, 1.1], [-1, 1.1], [-1, -0.1]]
    right_wall = [[0.95, -0.1], [0.95, 1.1], [2, 1.1], [2, -0.1]]
    divider = [[0.45, -1], [0.45, 0.3], [0.55, 0.3], [0.55, -1]]

    def state_initializer():
        walls = [
            sprite_lib.Sprite(shape=np.array(v), x=0, y=0, c0=128, c1=128,
                              c2=128)
            for v in [bottom_wall, left_wall, right_wall, divider]
        ]
        return collections.OrderedDict([
            ('walls', walls),
            ('balls', ball_generator(disjoint=True)),
            ('agent', []),
        ])

    collision = physics_lib.Collision(
        elasticity=0.6, symmetric=False, update_angle_vel=False,
        max_recursion_depth=2)
    physics = physics_lib.Physics(
        (collision, 'balls', ['balls', 'walls']),
        (physics_lib.DownGravity(g=-0.001), 'balls'),
        updates_per_env_step=20,
    )
    task = tasks.CompositeTask(timeout_steps=100)
    action_space = action_spaces.Grid(action_layers='agent')
    observer = observers.PILRenderer(image_size=image_size, anti_aliasing=1)
    return {
        'state_initializer': state_initializer,
        'physics': physics,
        'task': task,
        'action_space': action_space,
        'observers': {'image': observer},
        'game_rules': (),
    }
