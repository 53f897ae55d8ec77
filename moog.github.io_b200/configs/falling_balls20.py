"""falling_balls-20: BASELINE.json config 2 (SURVEY section 8d scene 2).

The scene of moog_demos/example_configs/falling_balls.py with `num_sprites=20`
and `scale=0.05` (at the shipped scale 0.1 the spawn box cannot hold 20
disjoint balls, falling_balls.py:44-52).  Unchanged otherwise: four walls
(floor, two sides, a divider), balls that collide asymmetrically
(elasticity 0.6, no spin, recursion depth 2) with each other and the walls
under DownGravity(-0.001) at 20 substeps, a 100-step timeout, a Grid action
space on an empty layer and the 64x64 PILRenderer with raw RGB colours.
"""

import collections

import numpy as np

from moog import action_spaces, observers, tasks
from moog import physics as phys
from moog import sprite as sprite_lib
from moog.state_initialization import distributions as distribs
from moog.state_initialization import sprite_generators

# absolute outlines of the static sprites (drawn with x = y = 0, scale 1)
_WALL_OUTLINES = collections.OrderedDict([
    ('floor', [[-1, 0.1], [2, 0.1], [2, -1], [-1, -1]]),
    ('left', [[0.05, -0.1], [0.05, 1.1], [-1, 1.1], [-1, -0.1]]),
    ('right', [[0.95, -0.1], [0.95, 1.1], [2, 1.1], [2, -0.1]]),
    ('divider', [[0.45, -1], [0.45, 0.3], [0.55, 0.3], [0.55, -1]]),
])
_GREY = dict(c0=128, c1=128, c2=128)
_BLUE = dict(c0=0, c1=0, c2=255)


def _ball_source(count, scale):
    spawn = [distribs.Continuous('x', 0.25, 0.75),
             distribs.Continuous('y', 0.5, 0.9),
             distribs.Continuous('x_vel', -0.01, 0.01)]
    return sprite_generators.generate_sprites(
        distribs.Product(spawn, scale=scale, shape='circle', mass=1., **_BLUE), num_sprites=count)


def _physics():
    bounce = phys.Collision(elasticity=0.6, symmetric=False, update_angle_vel=False, max_recursion_depth=2)
    return phys.Physics((bounce, 'balls', ['balls', 'walls']),
                        (phys.DownGravity(g=-0.001), 'balls'),
                        updates_per_env_step=20)


def get_config(level=None):
    """level: None or dict(num_sprites=..., scale=..., image_size=...)."""
    opts = dict(num_sprites=20, scale=0.05, image_size=(64, 64))
    opts.update(level or {})
    balls = _ball_source(opts['num_sprites'], opts['scale'])

    def state_initializer():
        walls = [sprite_lib.Sprite(shape=np.array(outline), x=0, y=0, **_GREY)
                 for outline in _WALL_OUTLINES.values()]
        return collections.OrderedDict(walls=walls, balls=balls(disjoint=True), agent=[])

    return dict(
        state_initializer=state_initializer,
        physics=_physics(),
        task=tasks.CompositeTask(timeout_steps=100),
        action_space=action_spaces.Grid(action_layers='agent'),
        observers=dict(image=observers.PILRenderer(image_size=opts['image_size'], anti_aliasing=1)),
        game_rules=(),
    )
