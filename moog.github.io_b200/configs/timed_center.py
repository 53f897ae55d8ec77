"""timed_center: a small first-person scene exercising the timing rules and
KeepNearCenter (moog/game_rules/timing.py, re_center.py) together with the
FirstPersonAgent renderer -- the rule set of the shipped first-person configs
(first_person_predators_prey, parallelogram_catch) without their 102-vertex
annulus sprites.  A circle agent (Joystick, Drag) moves among landmark polygons;
a cue square vanishes between steps 5 and 6 (TimedRule), the landmarks fade for
the first 8 steps (TemporaryRule) and turn red from step 12 on (DelayedRule).
"""

import collections

import numpy as np

from moog import action_spaces
from moog import game_rules
from moog import observers
from moog import physics as physics_lib
from moog import sprite
from moog import tasks
from moog.observers import polygon_modifiers


def get_config(level=None):
    del level

    def state_initializer():
        rng = np.random
        landmarks = [
            sprite.Sprite(x=0.2 + 0.15 * k + 0.02 * rng.rand(), y=0.25 + 0.5 * ((k * 7) % 5) / 5.,
                          shape=('triangle', 'square', 'pentagon', 'star_5')[k % 4], scale=0.07,
                          angle=0.3 * k, c0=0.1 + 0.2 * k, c1=1., c2=1., opacity=255)
            for k in range(5)]
        cue = [sprite.Sprite(x=0.5, y=0.85, shape='square', scale=0.1, c0=0.6, c1=1., c2=1.)]
        agent = [sprite.Sprite(x=0.5, y=0.5, shape='circle', scale=0.06, c0=0.33, c1=1., c2=0.66)]
        return collections.OrderedDict([('landmarks', landmarks), ('cue', cue), ('agent', agent)])

    physics = physics_lib.Physics(
        (physics_lib.Drag(coeff_friction=0.25), 'agent'), updates_per_env_step=5)

    def _fade(s):
        s.opacity = 128

    def _redden(s):
        s.c0 = 0.

    rules = (
        game_rules.KeepNearCenter('agent', ['landmarks', 'cue'], grid_x=0.1, grid_y=0.15),
        game_rules.TimedRule((5, 6), game_rules.VanishByFilter('cue')),
        game_rules.TemporaryRule(8, game_rules.ModifySprites('landmarks', _fade)),
        game_rules.DelayedRule(12, (game_rules.ModifySprites('landmarks', _redden),)),
    )

    return {
        'state_initializer': state_initializer,
        'physics': physics,
        'task': tasks.CompositeTask(timeout_steps=40),
        'action_space': action_spaces.Joystick(scaling_factor=0.02, action_layers='agent'),
        'observers': {'image': observers.PILRenderer(
            image_size=(64, 64), anti_aliasing=1, color_to_rgb='hsv_to_rgb',
            polygon_modifier=polygon_modifiers.FirstPersonAgent('agent'))},
        'game_rules': rules,
    }
