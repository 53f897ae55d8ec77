"""spawn_zoo: sprites that appear during the episode -- `CreateSprites` (game_rules/create_sprites.py:8-34)
behind random `ConditionalRule` conditions of the form `np.random.binomial(1, p)`, the pattern of
first_person_predators_prey.py:176-193.

  * `drops` appear one at a time anywhere on the field, with a random shape / size / colour, and must
    not overlap the walls, the agent or the drops already there (`without_overlapping` incl. the
    receiving layer itself); their generator gives up after a few tries (`fail_gracefully`), so a
    crowded field simply gets no new drop;
  * `sparks` appear two at a time on one of the four borders (a Mixture) with a velocity from a square
    annulus (a SetMinus), fly ballistically and vanish once they are far outside and moving away;
  * the agent eats drops (VanishOnContact + ContactReward).

The layers `drops` and `sparks` start empty, as `prey` / `predators` do in the shipped config.
"""

import collections

import numpy as np

from moog import action_spaces
from moog import game_rules
from moog import observers
from moog import physics as physics_lib
from moog import sprite
from moog import tasks
from moog.state_initialization import distributions as distribs
from moog.state_initialization import sprite_generators

# room the record gives the two layers that grow (compile_config(..., layer_capacity=LAYER_CAPACITY))
LAYER_CAPACITY = {'drops': 32, 'sparks': 32}


def _border(lo, hi):
    sides = [
        distribs.Product([distribs.Continuous('y', lo, hi)], x=lo),
        distribs.Product([distribs.Continuous('y', lo, hi)], x=hi),
        distribs.Product([distribs.Continuous('x', lo, hi)], y=lo),
        distribs.Product([distribs.Continuous('x', lo, hi)], y=hi),
    ]
    return distribs.Mixture(sides, probs=[0.1, 0.2, 0.3, 0.4])


def _ring(slow, fast):
    return distribs.SetMinus(
        distribs.Product([distribs.Continuous('x_vel', -fast, fast), distribs.Continuous('y_vel', -fast, fast)]),
        hold_out=distribs.Product([distribs.Continuous('x_vel', -slow, slow), distribs.Continuous('y_vel', -slow, slow)]))


def get_config(level=None):
    del level

    def state_initializer():
        walls = [
            sprite.Sprite(x=0.25, y=0.7, shape='square', scale=0.2, aspect_ratio=0.5, c0=0.6, c1=0.3, c2=0.6),
            sprite.Sprite(x=0.75, y=0.3, shape='square', scale=0.2, aspect_ratio=0.5, c0=0.6, c1=0.3, c2=0.6),
        ]
        agent = [sprite.Sprite(x=0.5 + 0.02 * np.random.rand(), y=0.5, shape='circle', scale=0.07, c0=0.33, c1=1., c2=1.)]
        return collections.OrderedDict([('walls', walls), ('drops', []), ('sparks', []), ('agent', agent)])

    drop_factors = distribs.Product(
        [distribs.Continuous('x', 0.05, 0.95), distribs.Continuous('y', 0.05, 0.95),
         distribs.Discrete('shape', ['triangle', 'square', 'circle', 'star_5']),
         distribs.Continuous('scale', 0.1, 0.17), distribs.Discrete('aspect_ratio', [1., 0.7]),
         distribs.Continuous('c0', 0., 1.)],
        c1=1., c2=1., mass=2.)
    new_drop = sprite_generators.generate_sprites(drop_factors, num_sprites=1, max_recursion_depth=6,
                                                  fail_gracefully=True)
    spark_factors = distribs.Product(
        [_border(-0.1, 1.1), _ring(0.02, 0.05), distribs.Continuous('scale', 0.02, 0.04)],
        shape='square', c0=0.12, c1=1., c2=1.)
    new_sparks = sprite_generators.generate_sprites(spark_factors, num_sprites=2)

    def _gone(s):
        return (s.x < -0.3 and s.x_vel < 0.) or (s.x > 1.3 and s.x_vel > 0.) or (
            s.y < -0.3 and s.y_vel < 0.) or (s.y > 1.3 and s.y_vel > 0.)

    rules = (
        game_rules.ConditionalRule(
            condition=lambda state: np.random.binomial(1, p=0.6),
            rules=game_rules.CreateSprites('drops', new_drop, without_overlapping=('walls', 'agent', 'drops'))),
        game_rules.ConditionalRule(
            condition=lambda state: np.random.binomial(1, 0.35),
            rules=game_rules.CreateSprites('sparks', new_sparks)),
        game_rules.VanishByFilter('sparks', _gone),
        game_rules.VanishOnContact(vanishing_layer='drops', contacting_layer='agent'),
    )

    physics = physics_lib.Physics(
        (physics_lib.Drag(coeff_friction=0.25), 'agent'),
        updates_per_env_step=4)

    task = tasks.CompositeTask(
        tasks.ContactReward(1., layers_0='agent', layers_1='drops'),
        timeout_steps=60)

    return {
        'state_initializer': state_initializer,
        'physics': physics,
        'task': task,
        'action_space': action_spaces.Joystick(scaling_factor=0.01, action_layers='agent'),
        'observers': {'image': observers.PILRenderer(image_size=(64, 64), anti_aliasing=1, color_to_rgb='hsv_to_rgb')},
        'game_rules': rules,
    }
