"""pacman-64: BASELINE.json config 4 (SURVEY section 8d scene 4).

The game of moog_demos/example_configs/pacman.py at level 0 (8x8 maze centred in
a 12x12 arena, two ghosts; level 1: 10x10, three ghosts) with the renderer at
64x64 -- BASELINE.json quotes the maze tasks at 64x64 and a 256x256 canvas does
not fit one CTA's shared memory.  Walls are one square per maze cell, a prey
pellet sits on every open cell, ghosts walk the maze at random
(RandomMazeWalk 0.015), MazePhysics keeps agent / prey / ghosts on the grid at
constant speed 0.015 (one substep), the agent eats pellets (+1), loses 5 on a
ghost (episode ends), the episode also ends 5 steps after the last pellet or
after 1000 steps, and pellets / ghosts are "unglued" (mass 1) once the agent
first moves.  Grid actions set the agent's velocity.
"""

import collections

import numpy as np

from moog import action_spaces, game_rules, maze_lib, observers, sprite, tasks
from moog import physics as phys

_SPEED = 0.015
_ARENA = 12
_LOOKS = dict(
    agent=dict(shape='circle', scale=0.05, c0=0.33, c1=1., c2=0.66),
    prey=dict(shape='circle', scale=0.025, c0=0.2, c1=1., c2=1.),
    ghost=dict(shape='circle', scale=0.05, mass=np.inf, c0=0., c1=1., c2=0.8),
    wall=dict(c0=0., c1=0., c2=0.8),
)
_LEVELS = {None: {}, 0: {}, 1: dict(num_ghosts=3, maze_size=10)}


def _at(maze, cell, look):
    """A sprite at the centre of maze cell (row, column)."""
    centre = maze.grid_side * (0.5 + np.array(cell))
    return sprite.Sprite(x=centre[1], y=centre[0], **look)


def _initializer(maze_size, num_ghosts):
    def state_initializer():
        cells = maze_lib.generate_random_maze_matrix(size=maze_size, ambient_size=_ARENA)
        maze = maze_lib.Maze(np.flip(cells, axis=0))
        walls = maze.to_sprites(**_LOOKS['wall'])
        spots = maze.sample_distinct_open_points(1 + num_ghosts)
        agent = [_at(maze, spots[0], _LOOKS['agent'])]
        ghosts = [_at(maze, spot, _LOOKS['ghost']) for spot in spots[1:]]
        prey = [_at(maze, cell, _LOOKS['prey']) for cell in np.argwhere(maze.maze == 0)]
        return collections.OrderedDict(walls=walls, prey=prey, ghosts=ghosts, agent=agent)
    return state_initializer


def _rules():
    def _unglue(s):
        s.mass = 1.

    def _agent_moves(state):
        return not np.all(state['agent'][0].velocity == 0)

    eat = game_rules.VanishOnContact(vanishing_layer='prey', contacting_layer='agent')
    unglue = game_rules.ConditionalRule(
        condition=_agent_moves, rules=game_rules.ModifySprites(('prey', 'ghosts'), _unglue))
    return (eat, unglue)


def get_config(level=None):
    """level: None / 0 / 1 or dict(num_ghosts=..., maze_size=..., image_size=...)."""
    opts = dict(num_ghosts=2, maze_size=8, image_size=(64, 64))
    opts.update(_LEVELS[level] if not isinstance(level, dict) else level)

    on_grid = phys.MazePhysics(maze_layer='walls', avatar_layers=('agent', 'prey', 'ghosts'),
                               constant_speed=_SPEED)
    physics = phys.Physics((phys.RandomMazeWalk(speed=_SPEED), ['ghosts']),
                           updates_per_env_step=1, corrective_physics=[on_grid])
    task = tasks.CompositeTask(
        tasks.ContactReward(-5, layers_0='agent', layers_1='ghosts', reset_steps_after_contact=0),
        tasks.ContactReward(1, layers_0='agent', layers_1='prey'),
        tasks.Reset(condition=lambda state: len(state['prey']) == 0, steps_after_condition=5),
        timeout_steps=1000)
    return dict(
        state_initializer=_initializer(opts['maze_size'], opts['num_ghosts']),
        physics=physics,
        task=task,
        action_space=action_spaces.Grid(scaling_factor=_SPEED, action_layers='agent',
                                        control_velocity=True, momentum=0.5),
        observers=dict(image=observers.PILRenderer(
            image_size=opts['image_size'], anti_aliasing=1, color_to_rgb='hsv_to_rgb')),
        game_rules=_rules(),
    )
