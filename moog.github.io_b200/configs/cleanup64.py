"""cleanup-64: BASELINE.json config 4 (SURVEY section 8d scene 4).

moog_demos/example_configs/cleanup.py (`get_config(None)`): three circle agents
with Drag and an asymmetric Collision(0.25) against 4 border walls (K = 5), a
2 x 6 grid of fountains and one of fruits whose HSV value says clean / ripe
(1.0) or bad (0.3), four contact-triggered rules -- agents touching a ripe
fruit poison one random clean fountain and spoil the fruit, agents touching a
bad fountain ripen one random fruit and clean the fountain --,
ContactReward(1) for agent_0 on ripe fruit, a Composite of three Joysticks,
64x64 HSV renderer (the RawState observer of the shipped file only feeds its
hand-written demo agents and is left out).
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
from moog.state_initialization import distributions as distribs
from moog.state_initialization import sprite_generators

_GOOD, _BAD, _THRESHOLD = 1., 0.3, 0.6
_AGENTS = ('agent_0', 'agent_1', 'agent_2')


def _grid(ys, **factors):
    gx, gy = np.meshgrid(np.linspace(0.1, 0.9, 6), np.linspace(ys[0], ys[1], 2))
    return [sprite.Sprite(x=x, y=y, **factors) for x, y in zip(np.ravel(gx), np.ravel(gy))]


def get_config(level=None):
    del level
    base = distribs.Product(
        [distribs.Continuous('x', 0., 1.), distribs.Continuous('y', 0.35, 0.65)],
        shape='circle', scale=0.1, c1=1., c2=0.7)
    generators = [sprite_generators.generate_sprites(distribs.Product([base], c0=c0), num_sprites=1)
                  for c0 in (0.2, 0.1, 0.)]
    walls = shapes.border_walls(visible_thickness=0.05, c0=0., c1=0., c2=0.5)
    fountains = _grid((0.75, 0.9), shape='circle', scale=0.05, c0=0.6, c1=1., c2=_BAD)
    fruits = _grid((0.1, 0.25), shape='circle', scale=0.05, c0=0.3, c1=1., c2=_BAD)

    def state_initializer():
        agents = [g(without_overlapping=walls) for g in generators]
        return collections.OrderedDict([
            ('walls', walls), ('fountains', fountains), ('fruits', fruits),
            ('agent_2', agents[2]), ('agent_1', agents[1]), ('agent_0', agents[0])])

    physics = physics_lib.Physics(
        (physics_lib.Drag(coeff_friction=0.25), list(_AGENTS)),
        (physics_lib.Collision(elasticity=0.25, symmetric=False), list(_AGENTS), 'walls'),
        updates_per_env_step=5)

    task = tasks.ContactReward(
        1, layers_0='agent_0', layers_1='fruits', condition=lambda s_0, s_1: s_1.c2 > _THRESHOLD)

    action_space = action_spaces.Composite(**{
        name: action_spaces.Joystick(scaling_factor=0.005, action_layers=name) for name in _AGENTS})

    def _make_bad(s):
        s.c2 = _BAD

    def _make_good(s):
        s.c2 = _GOOD

    # The state condition of cleanup.py:183-193.  The compiler recognises this helper by its
    # exact form (moog_b200/lambdas.py, MOOG_SC_CONTACT_ANY_COUNT): Python's `or` decides how
    # many overlap tests run, which is part of the reference's observable call sequence.
    def agents_contacting_layer(state, layer, value):
        n_contact = 0
        for s in state[layer]:
            if s.c2 != value:
                continue
            n_contact += (
                s.overlaps_sprite(state['agent_0'][0]) or
                s.overlaps_sprite(state['agent_1'][0]) or
                s.overlaps_sprite(state['agent_2'][0])
            )
        return n_contact

    poison_fountains = game_rules.ConditionalRule(
        condition=lambda s: agents_contacting_layer(s, 'fruits', _GOOD),
        rules=game_rules.ModifySprites(
            layers='fountains', modifier=_make_bad, sample_one=True, filter_fn=lambda s: s.c2 > _THRESHOLD))
    ripen_fruits = game_rules.ConditionalRule(
        condition=lambda s: agents_contacting_layer(s, 'fountains', _BAD),
        rules=game_rules.ModifySprites(
            layers='fruits', modifier=_make_good, sample_one=True, filter_fn=lambda s: s.c2 < _THRESHOLD))
    spoil_fruits = game_rules.ModifyOnContact(
        layers_0='fruits', layers_1=_AGENTS, modifier_0=_make_bad, filter_0=lambda s: s.c2 > _THRESHOLD)
    clean_fountains = game_rules.ModifyOnContact(
        layers_0='fountains', layers_1=_AGENTS, modifier_0=_make_good, filter_0=lambda s: s.c2 < _THRESHOLD)

    return {
        'state_initializer': state_initializer,
        'physics': physics,
        'task': task,
        'action_space': action_space,
        'observers': {'image': observers.PILRenderer(
            image_size=(64, 64), anti_aliasing=1, color_to_rgb='hsv_to_rgb')},
        'game_rules': (poison_fountains, spoil_fruits, ripen_fruits, clean_fountains),
    }
