"""The reference's own known-answer tests for the hot path, restated as data.

Source of every vector below (parameters and expected end states, 4 decimals,
tolerance `_ATOL = 0.001`):
    /root/reference/tests/moog/physics/test_collisions.py:98-293
        testCirclesSameMass (10), testCirclesDifferentMass (4), testTriangles (5)
    /root/reference/tests/moog/physics/test_tether_physics.py:109-216
        testTetherNoAngleUpdate, testTetherAngleUpdate, testTetherAnchor

The reference drives `Collision.step` directly over explicit sprite pairs
(`_apply_pairwise_force`, test_collisions.py:36-51) and then calls
`update_pos_from_vel(1.)`; that is exactly one `Physics.step` with
`updates_per_env_step=1` (moog/physics/physics.py:88-117) when the layers are
laid out so that `itertools.product` visits the pairs in the test's order:

    symmetric:  inds = (0,0) (1,0) (0,1) (1,1)  ->  one layer [sprite_1, sprite_0]
                paired with itself: (s1,s1) (s1,s0) (s0,s1) (s0,s0)
    asymmetric: inds = (1,0)                    ->  layers a=[sprite_1], b=[sprite_0]

`build(case)` returns (config, state) made of the MOOG-compatible host classes
of this repo, so the same vectors check the oracle (CPU) and the CUDA path.
"""
import collections

import numpy as np

ATOL = 0.001   # test_collisions.py:33, test_tether_physics.py:35

# (init_pos_0, init_vel_0, out_pos_0, out_vel_0, out_pos_1, out_vel_1, elasticity, symmetric)
CIRCLES_SAME_MASS = [
    ([0.5, 0.35], [0., 0.], [0.5, 0.35], [0., 0.], [0.5, 0.4827], [0., 0.01], 1., False),
    ([0.5, 0.35], [0., 0.], [0.5, 0.3287], [0., -0.01], [0.5, 0.4613], [0., 0.], 1., True),
    ([0.5, 0.35], [0., 0.], [0.5, 0.3337], [0., -0.0075], [0.5, 0.4563], [0., -0.0025], 0.5, True),
    ([0.5, 0.35], [0., 0.], [0.5, 0.3387], [0., -0.005], [0.5, 0.4513], [0., -0.005], 0., True),
    ([0.5, 0.35], [0., 0.01], [0.5, 0.3287], [0., -0.01], [0.5, 0.5213], [0., 0.01], 1., True),
    ([0.44, 0.37], [0., 0.], [0.44, 0.37], [0., 0.], [0.5217, 0.4699], [0.0095, 0.0031], 1., False),
    ([0.44, 0.37], [0., 0.], [0.4291, 0.3550], [-0.0048, -0.0065], [0.5109, 0.4550], [0.0048, -0.0035], 1., True),
    ([0.44, 0.37], [0., 0.], [0.4315, 0.3583], [-0.0036, -0.0049], [0.5085, 0.4517], [0.0036, -0.0051], 0.5, True),
    ([0.44, 0.37], [0., 0.01], [0.4006, 0.3758], [-0.0095, -0.0031], [0.5394, 0.4942], [0.0095, 0.0031], 1., True),
    ([0.43, 0.36], [0.015, 0.01], [0.4793, 0.3286], [0.0051, -0.0123], [0.5407, 0.5314], [0.0099, 0.0123], 1., True),
]

# (init_pos_0, init_vel_0, out_pos_0, out_vel_0, out_pos_1, out_vel_1); elasticity 1, symmetric, mass_1 = 2
CIRCLES_DIFFERENT_MASS = [
    ([0.5, 0.35], [0., 0.], [0.5, 0.3220], [0., -0.0133], [0.5, 0.4547], [0., -0.0033]),
    ([0.5, 0.35], [0., 0.], [0.5, 0.3220], [0., -0.0133], [0.5, 0.4547], [0., -0.0033]),
    ([0.44, 0.37], [0., 0.01], [0.3879, 0.3583], [-0.0127, -0.0075], [0.5267, 0.4768], [0.0063, -0.0013]),
    ([0.43, 0.36], [0.015, 0.01], [0.4661, 0.2989], [0.0018, -0.0197], [0.5275, 0.5017], [0.0066, 0.0048]),
]

# (init_angle_vel_0, out_pos_0, out_vel_0, out_angle_vel_0, out_pos_1, out_vel_1, out_angle_vel_1,
#  elasticity, update_angle_vel); symmetric, 10 steps
TRIANGLES = [
    (0., [0.5064, 0.6776], [-0.0044, 0.0024], 0., [0.6369, 0.5358], [0.0044, -0.0024], 0., 1., False),
    (0., [0.5411, 0.6689], [0.0025, 0.0006], -0.0911, [0.6022, 0.5444], [-0.0025, -0.0006], 0.0362, 1., True),
    (0., [0.5442, 0.6681], [0.0031, 0.0005], -0.0683, [0.5991, 0.5452], [-0.0031, -0.0005], 0.0271, 0.5, True),
    (0.1, [0.4950, 0.6804], [-0.0021, 0.0018], -0.1215, [0.6483, 0.5329], [0.0021, -0.0018], 0.0720, 1., True),
    (-0.02, [0.5486, 0.6670], [0.0035, 0.0004], -0.0800, [0.5947, 0.5463], [-0.0035, -0.0004], 0.0250, 1., True),
]

# name -> (tether kwargs, state after step 1, state after step 45); each state is
# [position, velocity, angle_vel] per sprite
TETHER = {
    'no_angle_update': (
        dict(update_angle_vel=False),
        [[[0.5133, 0.6933], [0.0133, -0.0067], 0.], [[0.2133, 0.5933], [0.0133, -0.0067], 0.],
         [[0.6133, 0.2933], [0.0133, -0.0067], 0.]],
        [[[0.7710, 0.4900], [-0.0005, 0.], 0.], [[0.4710, 0.3900], [-0.0005, 0.], 0.],
         [[0.8671, 0.0939], [-0.0005, 0.], 0.]]),
    'angle_update': (
        dict(update_angle_vel=True),
        [[[0.5206, 0.6900], [0.0205, -0.0103], -0.0447], [[0.2165, 0.6035], [0.0166, 0.0033], -0.0447],
         [[0.6027, 0.2860], [0.0025, -0.0140], -0.0447]],
        [[[0.8341, 0.3401], [-0.0028, -0.0046], -0.0229], [[0.7139, 0.6271], [0.0037, -0.0018], -0.0229],
         [[0.4545, 0.2062], [-0.0059, 0.0041], -0.0229]]),
    'anchor': (
        dict(anchor=np.array([0.2, 0.2])),
        [[[0.5190, 0.6881], [0.0188, -0.0122], -0.0385], [[0.2154, 0.5997], [0.0154, -0.0006], -0.0385],
         [[0.6036, 0.2845], [0.0033, -0.0155], -0.0385]],
        [[[0.6927, 0.5118], [0., 0.], 0.], [[0.3798, 0.5573], [0., 0.], 0.], [[0.6025, 0.1233], [0., 0.], 0.]]),
}


def _libs():
    import moog_b200  # noqa: F401  (puts the MOOG-compatible `moog` package on sys.path)
    from moog import action_spaces, observers, physics as physics_lib, shapes, sprite, tasks
    return action_spaces, observers, physics_lib, shapes, sprite, tasks


def _config(physics, state):
    action_spaces, _, _, _, _, tasks = _libs()
    layers = collections.OrderedDict((k, list(v)) for k, v in state.items())
    layers['agent'] = []
    return dict(state_initializer=lambda: layers, physics=physics,
                task=tasks.CompositeTask(timeout_steps=10 ** 6),
                action_space=action_spaces.Grid(action_layers='agent'),
                observers={}, game_rules=()), layers


def _pair_case(sprite_0, sprite_1, elasticity, symmetric, update_angle_vel, steps, expected):
    _, _, physics_lib, _, _, _ = _libs()
    force = physics_lib.Collision(elasticity=elasticity, symmetric=symmetric,
                                  update_angle_vel=update_angle_vel)
    if symmetric:
        state = collections.OrderedDict([('sprites', [sprite_1, sprite_0])])
        physics = physics_lib.Physics((force, 'sprites', 'sprites'), updates_per_env_step=1)
        where = {0: ('sprites', 1), 1: ('sprites', 0)}
    else:
        state = collections.OrderedDict([('a', [sprite_1]), ('b', [sprite_0])])
        physics = physics_lib.Physics((force, 'a', 'b'), updates_per_env_step=1)
        where = {0: ('b', 0), 1: ('a', 0)}
    config, layers = _config(physics, state)
    return dict(config=config, layers=layers, steps=steps, where=where, checks={steps: expected})


def collision_cases():
    """-> list of (name, case).  case['checks'][step] = {sprite index: (pos, vel, angle_vel|None)}."""
    _, _, _, _, sprite, _ = _libs()
    out = []
    for k, (p0, v0, op0, ov0, op1, ov1, el, sym) in enumerate(CIRCLES_SAME_MASS):
        s0 = sprite.Sprite(x=p0[0], y=p0[1], scale=0.1, shape='circle', x_vel=v0[0], y_vel=v0[1], c1=255)
        s1 = sprite.Sprite(x=0.5, y=0.5, scale=0.1, shape='circle', y_vel=-0.01, c0=255)
        out.append(('circles_same_mass_%d' % k,
                    _pair_case(s0, s1, el, sym, False, 6, {0: (op0, ov0, None), 1: (op1, ov1, None)})))
    for k, (p0, v0, op0, ov0, op1, ov1) in enumerate(CIRCLES_DIFFERENT_MASS):
        s0 = sprite.Sprite(x=p0[0], y=p0[1], scale=0.1, shape='circle', x_vel=v0[0], y_vel=v0[1], c1=255)
        s1 = sprite.Sprite(x=0.5, y=0.5, scale=0.1, shape='circle', y_vel=-0.01, c0=255, mass=2.)
        out.append(('circles_different_mass_%d' % k,
                    _pair_case(s0, s1, 1., True, False, 6, {0: (op0, ov0, None), 1: (op1, ov1, None)})))
    for k, (w0, op0, ov0, ow0, op1, ov1, ow1, el, upd) in enumerate(TRIANGLES):
        s0 = sprite.Sprite(x=0.5, y=0, scale=0.05, shape=np.array([[1, 1], [1, 3], [-2, -2]]),
                           x_vel=0.005, y_vel=0., c0=255, angle=1., angle_vel=w0)
        s1 = sprite.Sprite(x=0.31, y=0.88, scale=0.05, shape=np.array([[2, 1], [0, 1], [-1, -3]]),
                           x_vel=-0.005, y_vel=0., c1=255)
        out.append(('triangles_%d' % k,
                    _pair_case(s0, s1, el, True, upd, 10, {0: (op0, ov0, ow0), 1: (op1, ov1, ow1)})))
    return out


def tether_cases():
    _, _, physics_lib, shapes, sprite, _ = _libs()
    out = []
    for name, (kwargs, step_1, final) in TETHER.items():
        sprites = [
            sprite.Sprite(x=0.5, y=0.7, scale=0.1, shape='triangle', x_vel=0.04, y_vel=-0.02, c0=255, angle=2.),
            sprite.Sprite(x=0.2, y=0.6, scale=0.1, shape='triangle', x_vel=0., y_vel=0., c1=255, angle=1.),
            sprite.Sprite(x=0.6, y=0.3, scale=0.1, shape='triangle', x_vel=0., y_vel=0., c2=255),
        ]
        walls = shapes.border_walls(visible_thickness=0.05, c0=128, c1=128, c2=128)
        state = collections.OrderedDict([('walls', walls), ('sprites', sprites)])
        collision = physics_lib.Collision(elasticity=0., symmetric=False, update_angle_vel=True)
        physics = physics_lib.Physics((collision, 'sprites', 'walls'),
                                      corrective_physics=[physics_lib.Tether('sprites', **kwargs)],
                                      updates_per_env_step=10)
        config, layers = _config(physics, state)
        where = {i: ('sprites', i) for i in range(3)}
        checks = {1: {i: tuple(step_1[i]) for i in range(3)}, 45: {i: tuple(final[i]) for i in range(3)}}
        out.append(('tether_' + name, dict(config=config, layers=layers, steps=45, where=where, checks=checks)))
    return out


def all_cases():
    return collision_cases() + tether_cases()


def compile_case(case):
    """-> (program, packed state arrays of one env)."""
    from moog_b200 import compiler
    states = [case['layers']]
    prog = compiler.compile_config(case['config'], states)
    return prog, compiler.pack_states(prog, states)


def check_state(case, prog, dyn, step, name=''):
    """Compare the slots named by the case with the reference's known answers."""
    for idx, (pos, vel, w) in case['checks'][step].items():
        layer, k = case['where'][idx]
        s = prog.layer_off[prog.layer_index(layer)] + k
        got_pos, got_vel, got_w = dyn[0:2, s], dyn[2:4, s], dyn[5, s]
        assert np.allclose(got_pos, pos, atol=ATOL), (name, step, idx, 'position', got_pos, pos)
        assert np.allclose(got_vel, vel, atol=ATOL), (name, step, idx, 'velocity', got_vel, vel)
        if w is not None:
            assert np.allclose(got_w, w, atol=ATOL), (name, step, idx, 'angle_vel', got_w, w)


# ---------------------------------------------------------------------------
# rarely-hit Collision branches: values recorded from the reference itself
# (oracle/gen_kat_extra.py -> tests/golden/kat_rare_branches.json)
# ---------------------------------------------------------------------------
RARE_ATOL = 1e-12


def rare_branch_cases():
    import json
    import os
    _, _, physics_lib, _, sprite, _ = _libs()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'kat_rare_branches.json')
    with open(path) as f:
        data = json.load(f)
    out = []
    for name in sorted(data):
        case = data[name]

        def _sprite(kw):
            kw = dict(kw)
            if not isinstance(kw['shape'], str):
                kw['shape'] = np.array(kw['shape'])
            return sprite.Sprite(**kw)

        s0, s1 = _sprite(case['s0']), _sprite(case['s1'])
        force = physics_lib.Collision(**case['collision'])
        state = collections.OrderedDict([('a', [s0]), ('b', [s1])])
        physics = physics_lib.Physics((force, 'a', 'b'), updates_per_env_step=1)
        config, layers = _config(physics, state)
        out.append((name, dict(config=config, layers=layers, steps=1,
                               where={0: ('a', 0), 1: ('b', 0)}, expected=case['expected'])))
    return out


def check_rare(case, prog, dyn, name=''):
    exp = case['expected']
    for idx, (pk, vk) in enumerate((('pos0', 'vel0'), ('pos1', 'vel1'))):
        layer, k = case['where'][idx]
        s = prog.layer_off[prog.layer_index(layer)] + k
        assert np.allclose(dyn[0:2, s], exp[pk], rtol=0, atol=RARE_ATOL), (name, idx, 'position', dyn[0:2, s], exp[pk])
        assert np.allclose(dyn[2:4, s], exp[vk], rtol=0, atol=RARE_ATOL), (name, idx, 'velocity', dyn[2:4, s], exp[vk])


# ---------------------------------------------------------------------------
# Episode timing: /root/reference/tests/moog/env_wrappers/test_simulation.py:35-128
# ---------------------------------------------------------------------------
# testStep (:68-78): with these Grid actions the agent (0.1 per step, control_velocity) reaches
# the target at the 4th action; ContactReward(reset_steps_after_contact=2) then terminates the
# episode at the 6th and at no earlier step.  testSimStepSimPop (:80-128) replays prefixes of the
# same sequence through sim_step / sim_pop and expects the same timing after every restore.
SIM_ACTIONS = [1, 4, 3, 1, 2, 0]
SIM_INIT = [1, 4, 3]             # :82
SIM_POP_0 = [-1]                 # :83
SIM_REWARD_0 = [3, 1, 2, 0]      # :84
SIM_POP_1 = [-1, -2]             # :85
SIM_REWARD_1 = [1, 2, 0]         # :86


def simulation_timing_config():
    """get_env() of test_simulation.py:35-63 built from this repo's MOOG-compatible classes; the
    ModifyMetaState rule is left out (meta_state is host-side Python in MOOG and plays no part
    in the timing)."""
    action_spaces, observers, physics_lib, _, sprite, tasks = _libs()

    def _state_initializer():
        agent = sprite.Sprite(x=0.5, y=0.5, scale=0.1, c0=128)
        target = sprite.Sprite(x=0.75, y=0.5, scale=0.1, c1=128)
        return collections.OrderedDict([('agent', [agent]), ('target', [target])])

    return dict(
        state_initializer=_state_initializer,
        physics=physics_lib.Physics(),
        task=tasks.ContactReward(1., 'agent', 'target', reset_steps_after_contact=2),
        action_space=action_spaces.Grid(0.1, action_layers='agent', control_velocity=True),
        observers={'image': observers.PILRenderer(image_size=(64, 64))},
        game_rules=(),
    )
