"""pacman-64: BASELINE.json config 4 (SURVEY section 8d scene 4).

moog_demos/example_configs/pacman.py (`get_config(0)`: 8x8 maze in a 12x12
arena, 2 ghosts) with the PILRenderer at 64x64 -- BASELINE.json quotes the maze
tasks at 64x64; the shipped 256x256 canvas does not fit one CTA's shared
memory.  Everything else follows pacman.py:27-152: RandomMazeWalk(0.015) on the
ghosts, MazePhysics(constant_speed=0.015) on agent / prey / ghosts, K = 1,
VanishOnContact(prey, agent), the `unglue` ConditionalRule, two ContactRewards,
Reset 5 steps after the last prey, timeout 1000, Grid(control_velocity).
"""

import collections

import numpy as np

from moog import action_spaces
from moog import game_rules
from moog import maze_lib
from moog import observers
from moog import physics as physics_lib
from moog import sprite
from moog import tasks


def get_config(level=None):
    """level: None / 0 / 1 or dict(num_ghosts=..., maze_size=..., image_size=...)."""
    if level in (None, 0):
        level = {}
    elif level == 1:
        level = dict(num_ghosts=3, maze_size=10)
    num_ghosts = level.get('num_ghosts', 2)
    maze_size = level.get('maze_size', 8)
    image_size = level.get('image_size', (64, 64))

    agent_factors = dict(shape='circle', scale=0.05, c0=0.33, c1=1., c2=0.66)
    prey_factors = dict(shape='circle', scale=0.025, c0=0.2, c1=1., c2=1.)
    ghost_factors = dict(shape='circle', scale=0.05, mass=np.inf, c0=0., c1=1., c2=0.8)

    def state_initializer():
        maze = maze_lib.generate_random_maze_matrix(size=maze_size, ambient_size=12)
        maze = maze_lib.Maze(np.flip(maze, axis=0))
        walls = maze.to_sprites(c0=0., c1=0., c2=0.8)
        points = maze.sample_distinct_open_points(1 + num_ghosts)
        positions = [maze.grid_side * (0.5 + np.array(x)) for x in points]
        agent = [sprite.Sprite(x=positions[0][1], y=positions[0][0], **agent_factors)]
        ghosts = [sprite.Sprite(x=p[1], y=p[0], **ghost_factors) for p in positions[1:]]
        prey = []
        for p in np.argwhere(maze.maze == 0):
            pos = maze.grid_side * (0.5 + np.array(p))
            prey.append(sprite.Sprite(x=pos[1], y=pos[0], **prey_factors))
        return collections.OrderedDict([
            ('walls', walls), ('prey', prey), ('ghosts', ghosts), ('agent', agent)])

    maze_physics = physics_lib.MazePhysics(
        maze_layer='walls', avatar_layers=('agent', 'prey', 'ghosts'), constant_speed=0.015)
    physics = physics_lib.Physics(
        (physics_lib.RandomMazeWalk(speed=0.015), ['ghosts']),
        updates_per_env_step=1, corrective_physics=[maze_physics])

    ghost_task = tasks.ContactReward(
        -5, layers_0='agent', layers_1='ghosts', reset_steps_after_contact=0)
    prey_task = tasks.ContactReward(1, layers_0='agent', layers_1='prey')
    reset_task = tasks.Reset(
        condition=lambda state: len(state['prey']) == 0, steps_after_condition=5)
    task = tasks.CompositeTask(ghost_task, prey_task, reset_task, timeout_steps=1000)

    action_space = action_spaces.Grid(
        scaling_factor=0.015, action_layers='agent', control_velocity=True, momentum=0.5)

    observer = observers.PILRenderer(
        image_size=image_size, anti_aliasing=1, color_to_rgb='hsv_to_rgb')

    def _unglue(s):
        s.mass = 1.

    def _unglue_condition(state):
        return not np.all(state['agent'][0].velocity == 0)

    unglue = game_rules.ConditionalRule(
        condition=_unglue_condition,
        rules=game_rules.ModifySprites(('prey', 'ghosts'), _unglue))
    vanish_on_contact = game_rules.VanishOnContact(
        vanishing_layer='prey', contacting_layer='agent')

    return {
        'state_initializer': state_initializer,
        'physics': physics,
        'task': task,
        'action_space': action_space,
        'observers': {'image': observer},
        'game_rules': (vanish_on_contact, unglue),
    }
