"""Numeric `sprite.metadata[...]` on the device path and config callables that branch on traced
values (`if` / `elif` / `else`): /root/reference/moog_demos/example_configs/red_green.py:144-176
(`agent.metadata = {'true_contact_color': ...}` read by the ContactReward's reward_fn) and
bounce_box_contact_prediction.py:94-110.  The reference trajectories are the `red_green` and
`predict_zoo` goldens (tests/util.SCENES); here: every branch, the lowering itself, metadata that
moves with its sprite, and the host-side `Physics.step` of the state initializers."""
import collections

import numpy as np
import pytest

from tests import util


def _answer(agent, box):
    said_right = box.x > 0.5
    if agent.metadata['goal'] == said_right:
        return 1.
    elif agent.metadata['when'] > 30:
        return -0.5
    else:
        return -1.


def _config(reward_fn, metadata, rules=()):
    import moog_b200  # noqa: F401
    from moog import action_spaces, physics as physics_lib, sprite, tasks

    def state_initializer(md=None):
        boxes = [sprite.Sprite(x=0.4, y=0.5, shape='square', scale=0.1, c0=1.),
                 sprite.Sprite(x=0.6, y=0.5, shape='square', scale=0.1, c0=2.)]
        agent = sprite.Sprite(x=0.5, y=0.5, shape='square', scale=0.15)      # touches both boxes
        agent.metadata = dict(md if md is not None else metadata[0])
        extra = [sprite.Sprite(x=0.1 + 0.2 * k, y=0.1, shape='triangle', scale=0.05, metadata={'goal': 10 + k, 'when': k})
                 for k in range(4)]
        return collections.OrderedDict([('boxes', boxes), ('extra', extra), ('agent', [agent])])

    cfg = dict(state_initializer=state_initializer, physics=physics_lib.Physics(updates_per_env_step=1),
               task=tasks.CompositeTask(tasks.ContactReward(reward_fn=reward_fn, layers_0='agent', layers_1='boxes'),
                                        timeout_steps=100),
               action_space=action_spaces.Grid(scaling_factor=0.01, action_layers='agent'), observers={}, game_rules=rules)
    return cfg, [state_initializer(md) for md in metadata]


def test_reward_branches_on_metadata_match_python():
    """The lowered reward (selects over both sides of every `if`) against the Python function itself
    called on the host sprites, for every combination of the metadata the branches read."""
    from moog_b200 import compiler, lambdas
    from oracle.oracle import Oracle
    metadata = [{'goal': g, 'when': w} for g in (0, 1) for w in (5, 31, 60)]
    cfg, states = _config(_answer, metadata)
    prog = compiler.compile_config(cfg, states)
    assert prog.meta_keys == ['goal', 'when'] and prog.header[compiler.H_N_META] == 2
    assert any(op == lambdas.X_SELECT for op, _, _ in prog.expr)
    arrays = compiler.pack_states(prog, states)
    orc = Oracle(prog, arrays)
    orc.post_reset()
    reward, _ = orc.step(np.full((len(states), 1), 4.0))
    want = [_answer(st['agent'][0], st['boxes'][-1]) for st in states]      # contact_reward.py:88-95: the last touching pair's reward
    assert reward.tolist() == want and len(set(want)) >= 3
    # a sprite without the key: NaN in the column (the reference raises KeyError there)
    S = prog.n_slots
    col = arrays['envf'][0, prog.meta_off:prog.meta_off + 2 * S].reshape(2, S)
    assert np.isnan(col[:, prog.layer_off[0]:prog.layer_off[0] + 2]).all()
    assert col[0, prog.layer_off[1]:prog.layer_off[1] + 4].tolist() == [10, 11, 12, 13]


def test_metadata_moves_with_its_sprite():
    """VanishByFilter pops the second `extra` sprite (vanish.py:31-39): the sprites behind it move down
    one slot and their metadata columns with them (read back through a reward that uses them)."""
    import moog_b200  # noqa: F401
    from moog import game_rules, tasks
    from moog_b200 import compiler
    from oracle.oracle import Oracle
    cfg, states = _config(_answer, [{'goal': 1, 'when': 2}],
                          rules=(game_rules.VanishByFilter('extra', lambda s: s.metadata['goal'] == 11),))
    cfg['task'] = tasks.CompositeTask(
        tasks.ContactReward(reward_fn=lambda a, e: e.metadata['goal'] + 100 * e.metadata['when'], layers_0='extra',
                            layers_1='extra', condition=lambda a, e: a.x < e.x), timeout_steps=100)
    prog = compiler.compile_config(cfg, states)
    orc = Oracle(prog, compiler.pack_states(prog, states))
    orc.post_reset()                       # the rules pass of reset() pops the sprite with goal == 11
    S, lo = prog.n_slots, prog.layer_off[1]
    assert orc.cnt[0, 1] == 3
    col = orc.envf[0, prog.meta_off:prog.meta_off + 2 * S].reshape(2, S)
    assert col[0, lo:lo + 3].tolist() == [10, 12, 13] and col[1, lo:lo + 3].tolist() == [0, 2, 3]


def test_branch_lowering_forms():
    """if / elif / else with returns, a conditional expression, nested ifs and a branch on an `and`;
    callables with effects (modifiers) may not branch on traced values."""
    from moog_b200 import lambdas

    def nested(a, b):
        if a.x > 0.5:
            if b.y > 0.5:
                return 1.
            return 2.
        return 3. if b.c0 > a.c0 else b.mass

    def both(a, b):
        if a.x > 0.5 and b.x > 0.5:
            return a.scale
        return -1.

    for fn in (nested, both):
        const, code = lambdas.pair_reward(fn)
        assert const == 0.0 and any(op == lambdas.X_SELECT for op, _, _ in code)

    def modifier(s):
        if s.x > 0.5:
            s.c0 = 1.
    with pytest.raises(lambdas.LoweringError):
        lambdas.compile_modifier(modifier)


@pytest.mark.gpu
def test_host_physics_step_predicts_what_the_device_then_does():
    """predict_zoo through the public API: the initializer rolls the puck forward with
    `physics.step(state)` on the host (a batch of one env through the CUDA physics kernel) and stores the
    goal it will reach and when in `agent.metadata`; the batched environment then steps the same initial
    states: the puck must reach THAT goal at THAT step, and the metadata-branching reward must pay what
    the Python function says."""
    import torch
    import moog_b200  # noqa: F401
    from moog_b200.batched_env import BatchedEnvironment
    from moog_b200.configs import predict_zoo
    cfg = predict_zoo.get_config()
    np.random.seed(7)
    states = [cfg['state_initializer']() for _ in range(6)]
    predicted = [(st['agent'][0].metadata['goal'], st['agent'][0].metadata['when']) for st in states]
    assert all(predict_zoo.STEP_RANGE[0] <= w < predict_zoo.STEP_RANGE[1] for _, w in predicted)
    N = len(states)
    env = BatchedEnvironment(**cfg, num_envs=N, device='cuda:0', seed=1, initial_states=states)
    prog, eng = env.program, env.engine
    eng.state.upload({k: v for k, v in moog_b200.compiler.pack_states(prog, states).items() if k in util.STATE_KEYS})
    eng.post_reset()
    env._started = True
    puck, goals = prog.layer_off[2], prog.layer_off[1]
    stopped_at = [None] * N
    for t in range(predict_zoo.STEP_RANGE[1] + 2):
        env.step(torch.full((N, 1), 4.0, dtype=torch.float64))
        touch = eng.overlap_pairs('puck', 'goals').cpu().numpy()[:, 0, :]
        for e in range(N):
            if stopped_at[e] is None and touch[e].any():
                stopped_at[e] = (int(np.argmax(touch[e])), t + 1)
    assert stopped_at == predicted, (stopped_at, predicted)
    st = eng.state.download()
    assert (st['dyn'][:, 2:4, puck] == 0).all(), 'the rule stopped every puck on its goal'
    # the agents answer: env e walks left when e is even, right when odd; rewards per the Python function
    paid = np.zeros(N)
    for t in range(12):
        act = torch.tensor([[0.0] if e % 2 == 0 else [1.0] for e in range(N)], dtype=torch.float64)
        ts = env.step(act)
        r = ts.reward.cpu().numpy()
        paid += np.where(np.isnan(r), 0.0, r)
    for e, (goal, when) in enumerate(predicted):
        said_right = e % 2 == 1
        want = 1. if goal == said_right else (-0.5 if when > 30 else -1.)
        assert paid[e] != 0 and np.sign(paid[e]) == np.sign(want) and abs(paid[e] / want - round(paid[e] / want)) < 1e-9, (e, paid[e], want)


@pytest.mark.reference
@pytest.mark.parametrize('module,level,key', [('red_green', 1, 'true_contact_color'),
                                              ('bounce_box_contact_prediction', False, 'will_contact')])
def test_shipped_prediction_configs_compile_unchanged(monkeypatch, module, level, key):
    """Build container only: moog_demos/example_configs/red_green.py and bounce_box_contact_prediction.py
    as shipped, on this repo's `moog` package -- the custom RadialVelocity distribution, initializers
    that roll `physics.step(state)` forward (served here by the CPU oracle as a TEST stand-in for the
    CUDA call of moog_b200/host_physics.py, which has its own GPU test), `agent.metadata`, the
    branching reward / the decision-tree Reset task -- compile and run; the metadata column holds what
    the initializer predicted."""
    import importlib
    import sys
    import moog_b200  # noqa: F401
    from moog_b200 import compiler, host_physics
    from oracle.oracle import Oracle

    def _oracle_step(physics, state):
        from moog import action_spaces, tasks
        cfg = dict(physics=physics, task=tasks.CompositeTask(), action_space=action_spaces.Grid(action_layers=()),
                   observers={}, game_rules=())
        prog = compiler.compile_config(cfg, [state])
        orc = Oracle(prog, compiler.pack_states(prog, [state]))
        orc.physics_step()
        for l, name in enumerate(prog.layer_names):
            for k, sp in enumerate(state[name]):
                s = prog.layer_off[l] + k
                sp.position = np.array(orc.dyn[0, 0:2, s])
                sp.velocity = np.array(orc.dyn[0, 2:4, s])

    monkeypatch.setattr(host_physics, 'step', _oracle_step)
    import moog  # noqa: F401  (this repo's package, before the reference's directory joins sys.path)
    monkeypatch.setattr(sys, 'path', sys.path + ['/root/reference'])
    for name in [m for m in sys.modules if m.startswith('moog_demos')]:
        monkeypatch.delitem(sys.modules, name)
    shipped = importlib.import_module('moog_demos.example_configs.' + module)
    np.random.seed(3)
    cfg = shipped.get_config(level)
    states = [cfg['state_initializer']() for _ in range(2)]
    prog = compiler.compile_config(cfg, states)
    assert prog.meta_keys == [key]
    arrays = compiler.pack_states(prog, states)
    agent = prog.layer_off[prog.layer_index('agent')]
    col = arrays['envf'][:, prog.meta_off + agent]
    assert col.tolist() == [float(st['agent'][0].metadata[key]) for st in states]
    orc = Oracle(prog, arrays)
    orc.post_reset()
    for _ in range(20):
        orc.step(np.full((2, 1), 4.0))
    assert (orc.envi[:, 2] == 0).all()


def _verdict(state):
    """bounce_box_contact_prediction.py:94-103 in structure: single sprites picked out of the state,
    overlap tests made lazily, metadata read only on the branch that needs it."""
    agent = state['agent'][0]
    if agent.overlaps_sprite(state['boxes'][0]):
        return -1 if agent.metadata['goal'] else 1
    elif agent.overlaps_sprite(state['boxes'][1]):
        return 2 if agent.metadata['goal'] else -2
    else:
        return 0


def test_state_decision_tree_matches_python_call_for_call():
    """Reset(condition=lambda state: f(state) != 0, reward_fn=f) with f a branching function over picked
    sprites: the MOOG_SC_TREE ops pay what Python pays, reset when Python resets, and make exactly the
    overlap calls Python makes (the second test only when the first one failed; the reward function
    only when the condition held)."""
    import moog_b200  # noqa: F401
    from moog import action_spaces, physics as physics_lib, sprite, tasks
    from moog_b200 import compiler
    from oracle.oracle import Oracle

    def state_initializer(x, goal):
        boxes = [sprite.Sprite(x=0.3, y=0.5, shape='square', scale=0.1), sprite.Sprite(x=0.7, y=0.5, shape='square', scale=0.1)]
        agent = sprite.Sprite(x=x, y=0.5, shape='circle', scale=0.08, metadata={'goal': goal})
        return collections.OrderedDict([('boxes', boxes), ('agent', [agent])])

    cases = [(x, goal) for x in (0.3, 0.5, 0.7) for goal in (False, True)]
    states = [state_initializer(x, goal) for x, goal in cases]
    cfg = dict(state_initializer=lambda: state_initializer(0.5, True), physics=physics_lib.Physics(updates_per_env_step=1),
               task=tasks.CompositeTask(tasks.Reset(condition=lambda state: _verdict(state) != 0, reward_fn=_verdict,
                                                    steps_after_condition=2), timeout_steps=100),
               action_space=action_spaces.Grid(scaling_factor=0.01, action_layers='agent'), observers={}, game_rules=())
    prog = compiler.compile_config(cfg, states)
    assert [o['kind'] for o in prog.ops].count(compiler.SC_TREE) == 2
    orc = Oracle(prog, compiler.pack_states(prog, states))
    orc.post_reset()
    reward, step_type = orc.step(np.full((len(states), 1), 4.0))
    want = [float(_verdict(st)) for st in states]
    assert reward.tolist() == want == [-0.0 + 1, -1, 0, 0, -2, 2]
    # overlap calls: left box only (1) when it is hit -- condition + reward function: 2 calls;
    # right box: 2 tests each time -> 4; nothing hit: the condition alone -> 2
    assert orc.counters[:, 0].tolist() == [2, 2, 2, 2, 4, 4]
    for _ in range(2):
        reward, step_type = orc.step(np.full((len(states), 1), 4.0))
    assert step_type.tolist() == [2, 2, 1, 1, 2, 2]          # steps_after_condition = 2
    # an index beyond the layer: the reference's IndexError becomes MOOG_ERR_BAD_INDEX
    bad = dict(cfg, task=tasks.CompositeTask(tasks.Reset(condition=lambda state: state['boxes'][5].x > 0, steps_after_condition=2)))
    prog2 = compiler.compile_config(bad, states)
    orc2 = Oracle(prog2, compiler.pack_states(prog2, states))
    orc2.post_reset()
    orc2.step(np.full((len(states), 1), 4.0))
    assert ((orc2.envi[:, 2] & 128) != 0).all()


class _Tagger(object):
    """A rule class of a config's own, in the style of functional_maze.py:18-67: numeric attributes of its
    own (one of them never touched by reset), a loop over a layer whose length varies, overlap tests,
    assignments to sprites under branches."""

    def __init__(self, threshold):
        self._threshold = threshold
        self.total = 0

    def reset(self, state, meta_state):
        del state, meta_state
        self.hits = 0

    def step(self, state, meta_state):
        del meta_state
        agent = state['agent'][0]
        for s in state['items']:
            if agent.overlaps_sprite(s):
                self.hits += 1
                self.total = self.total + 1
                s.c0 = s.c0 + 1.
        if self.hits > self._threshold:
            agent.mass = agent.mass * 2
            self.hits = 0


def test_user_defined_rule_is_traced_path_by_path():
    """lambdas.trace_rule: the rule's step() as a decision tree (MOOG_R_TREE) against the rule itself run
    in Python on host sprites, step by step, for states with 0..3 items of which some touch the agent."""
    import moog_b200  # noqa: F401
    from moog import action_spaces, physics as physics_lib, sprite, tasks
    from moog_b200 import compiler
    from oracle.oracle import Oracle

    def state_initializer(n_items):
        agent = sprite.Sprite(x=0.5, y=0.5, shape='square', scale=0.2, mass=1.)
        xs = [0.45, 0.9, 0.55]           # the first and the third touch the agent
        items = [sprite.Sprite(x=xs[k], y=0.5, shape='circle', scale=0.05, c0=float(k)) for k in range(n_items)]
        return collections.OrderedDict([('items', items), ('agent', [agent])])

    states = [state_initializer(n) for n in (3, 0, 1, 2, 3)]
    rule = _Tagger(threshold=3)
    cfg = dict(state_initializer=lambda: state_initializer(3), physics=physics_lib.Physics(updates_per_env_step=1),
               task=tasks.CompositeTask(timeout_steps=100), action_space=action_spaces.Grid(action_layers=()),
               observers={}, game_rules=(rule,))
    prog = compiler.compile_config(cfg, states)
    tree = [o for o in prog.ops if o['kind'] == compiler.R_TREE]
    assert len(tree) == 1 and tree[0]['i'][3] == 2          # two state variables: hits (reset) and total (constructor)
    assert prog.dpool[tree[0]['i'][4]:tree[0]['i'][4] + 2] == [0.0, 0.0]
    orc = Oracle(prog, compiler.pack_states(prog, states))
    orc.post_reset()                                   # reset(), then every rule is stepped once
    twins = [(_Tagger(threshold=3), st) for st in states]
    for r, st in twins:
        r.reset(st, None)
        r.step(st, None)
    lo, ag = prog.layer_off[0], prog.layer_off[1]
    base = tree[0]['i'][2]
    for t in range(7):
        for e, (r, st) in enumerate(twins):
            n = len(st['items'])
            assert orc.stat[e, 6, lo:lo + n].tolist() == [s.c0 for s in st['items']], (t, e, 'c0')
            assert orc.stat[e, 0, ag] == st['agent'][0].mass, (t, e, 'mass')
            assert orc.envf[e, base:base + 2].tolist() == [float(r.hits), float(r.total)], (t, e, 'rule attributes')
        orc.step(np.full((len(states), 1), 4.0))
        for r, st in twins:
            r.step(st, None)
    assert orc.stat[0, 0, ag] == 16.0 and orc.stat[1, 0, ag] == 1.0         # 2 hits per pass, 8 passes: doubled 4 times; no items: never
    assert (orc.envi[:, 2] == 0).all()


def test_rules_that_cannot_be_traced_are_refused():
    import moog_b200  # noqa: F401
    from moog import action_spaces, physics as physics_lib, sprite, tasks
    from moog_b200 import compiler

    def state_initializer():
        return collections.OrderedDict([('agent', [sprite.Sprite(x=0.5, y=0.5)]), ('items', [sprite.Sprite(x=0.2, y=0.2)])])

    class Random(object):
        def reset(self, state, meta_state):
            pass

        def step(self, state, meta_state):
            state['agent'][0].c0 = np.random.normal()        # (uniform / randint scalars ARE lowered: rule-noise columns)

    class MovesThenTests(object):
        def reset(self, state, meta_state):
            pass

        def step(self, state, meta_state):
            agent = state['agent'][0]
            agent.position = np.array([0.1, 0.1])
            if agent.overlaps_sprite(state['items'][0]):
                agent.c0 = 1.

    class ReadsItsOwnWrite(object):
        def reset(self, state, meta_state):
            pass

        def step(self, state, meta_state):
            agent = state['agent'][0]
            if agent.x > 0.2:
                agent.mass = 2.
                state['items'][0].mass = agent.mass * state['items'][0].mass     # fine: the traced value of agent.mass is 2
                agent.c0 = agent.c1
                agent.c1 = 5.
                agent.c2 = agent.c1 + state['items'][0].mass                      # reads items[0].mass, assigned above

    base = dict(state_initializer=state_initializer, physics=physics_lib.Physics(updates_per_env_step=1),
                task=tasks.CompositeTask(timeout_steps=10), action_space=action_spaces.Grid(action_layers=()), observers={})
    states = [state_initializer()]
    for cls, text in ((Random, 'random'), (MovesThenTests, 'moved'), (ReadsItsOwnWrite, 'earlier assignment')):
        with pytest.raises(compiler.CompileError, match=text):
            compiler.compile_config(dict(base, game_rules=(cls(),)), states)


def _phase_config(rules, meta=None, task=None):
    import moog_b200  # noqa: F401
    from moog import action_spaces, physics as physics_lib, sprite, tasks

    def state_initializer():
        agent = sprite.Sprite(x=0.5, y=0.5, shape='square', scale=0.1, c0=0.)
        cross = sprite.Sprite(x=0.52, y=0.5, shape='square', scale=0.05)
        return collections.OrderedDict([('cross', [cross]), ('agent', [agent])])

    cfg = dict(state_initializer=state_initializer, physics=physics_lib.Physics(updates_per_env_step=1),
               task=task or tasks.CompositeTask(timeout_steps=100),
               action_space=action_spaces.SetPosition(action_layers='agent'), observers={}, game_rules=rules,
               meta_state_initializer=meta)
    return cfg, [state_initializer()]


def test_phase_sequence_fixation_and_meta_state():
    """task_phases.py:18-141, fixation.py:19-54 and a dict meta_state, against what the reference's
    classes do (restated inline): one-time rules on a phase's first step only, continual rules every
    step, a phase ends by duration or by an end condition that reads meta_state, the sequence steps only
    the phase that was current when the pass began and names it in meta_state; a Reset task reads the
    name; stepping past the last phase is the reference's IndexError."""
    import moog_b200  # noqa: F401
    from moog import game_rules as gr, tasks
    from moog_b200 import compiler
    from oracle.oracle import Oracle

    def bump(s):
        s.c0 = s.c0 + 1.

    def mark(s):
        s.c1 = s.c1 + 10.

    phases = gr.PhaseSequence(
        gr.Phase(continual_rules=gr.Fixation('agent', 'cross', 0.1, 'held'),
                 end_condition=lambda state, meta_state: meta_state['held'] >= 3, name='fixate'),
        gr.Phase(one_time_rules=gr.ModifySprites('agent', mark), continual_rules=gr.ModifySprites('agent', bump),
                 duration=4, name='count'),
        gr.Phase(one_time_rules=gr.ModifySprites('agent', mark), duration=lambda: np.random.randint(2, 5), name='wait'),
        gr.Phase(name='done'),
        meta_state_phase_name_key='phase')
    task = tasks.CompositeTask(tasks.Reset(condition=lambda state, meta_state: meta_state['phase'] == 'done',
                                           reward_fn=lambda _: 7., steps_after_condition=1), timeout_steps=100)
    cfg, states = _phase_config((phases,), meta=lambda: {'phase': '', 'held': 0}, task=task)
    prog = compiler.compile_config(cfg, states)
    assert list(prog.meta_vars) == ['phase', 'held'] and prog.duration_draws == [(0, 2, 5)]
    kinds = [o['kind'] for o in prog.ops]
    assert kinds.count(compiler.R_PHASE_BEGIN) == 4 and kinds.count(compiler.R_PHASESEQ_BEGIN) == 1
    for wait in (2, 3, 4):
        orc = Oracle(prog, compiler.pack_states(prog, states))
        orc.post_reset(rule_noise=np.array([[(wait - 2 + 0.5) / 3]]))     # the reset pass is the first step of 'fixate'
        slot = prog.meta_vars
        ag = prog.layer_off[1]
        names = {float(prog.strings.index(n) + 1): n for n in prog.strings}
        trace, paid = [], 0.0
        for t in range(14):
            trace.append((names[orc.envf[0, slot['phase']]], orc.envf[0, slot['held']], orc.stat[0, 6, ag], orc.stat[0, 7, ag]))
            act = np.array([[0.5, 0.5]]) if t != 0 else np.array([[0.9, 0.9]])     # looks away once (seen by the next pass): the count restarts
            reward, step_type = orc.step(act)
            paid += reward[0]
            if step_type[0] == 2:
                break
        phases_seen = [p for p, _, _, _ in trace]
        # held: 1 after the reset pass, 2, 0 (looked away), 1, 2, 3 -> 'fixate' ends on the 6th pass
        assert [h for _, h, _, _ in trace[:6]] == [1, 2, 0, 1, 2, 3]
        assert phases_seen[:5] == ['fixate'] * 5 and phases_seen[5:9] == ['count'] * 4
        assert phases_seen[9:9 + wait] == ['wait'] * wait, (wait, phases_seen)
        assert trace[9][2:] == (4.0, 10.0)                 # bump every step of 'count', mark once
        assert trace[-1][2:] == (4.0, 20.0) and trace[-1][0] == 'done'
        assert paid == 7.0 and step_type[0] == 2 and orc.envi[0, 2] == 0       # paid once, when 'done' is first seen
    # past the last phase: IndexError in the reference
    short = gr.PhaseSequence(gr.Phase(duration=1, name='a'), gr.Phase(duration=1, name='b'))
    cfg, states = _phase_config((short,))
    prog = compiler.compile_config(cfg, states)
    orc = Oracle(prog, compiler.pack_states(prog, states))
    orc.post_reset()
    orc.step(np.array([[0.5, 0.5]]))
    assert orc.envi[0, 2] & 128
    # meta_state must be a dict of numbers / strings
    cfg, states = _phase_config((), meta=lambda: [1, 2])
    with pytest.raises(compiler.CompileError, match='dict'):
        compiler.compile_config(cfg, states)


@pytest.mark.reference
def test_shipped_multi_tracking_compiles_unchanged(monkeypatch):
    """Build container only: moog_demos/example_configs/multi_tracking_with_feature.py as shipped, on this
    repo's `moog` package (PhaseSequence of five phases, two Fixation rules, its own ChangeTargetFeature
    rule, dict meta_state, RawState observer ignored)."""
    import importlib
    import sys
    import moog_b200  # noqa: F401
    from moog_b200 import compiler
    from oracle.oracle import Oracle
    import moog  # noqa: F401  (this repo's package, before the reference's directory joins sys.path)
    monkeypatch.setattr(sys, 'path', sys.path + ['/root/reference'])
    for name in [m for m in sys.modules if m.startswith('moog_demos')]:
        monkeypatch.delitem(sys.modules, name)
    shipped = importlib.import_module('moog_demos.example_configs.multi_tracking_with_feature')
    np.random.seed(5)
    cfg = shipped.get_config(3)
    states = [cfg['state_initializer']() for _ in range(2)]
    prog = compiler.compile_config(cfg, states)
    assert list(prog.meta_vars) == ['phase', 'fixation_duration', 'response_duration'] and len(prog.duration_draws) == 1
    orc = Oracle(prog, compiler.pack_states(prog, states))
    Oracle.set_seed(3)
    orc.post_reset()
    for _ in range(30):
        orc.step(np.full((2, 2), 0.5))
    assert (orc.envi[:, 2] == 0).all()
    assert (orc.envf[:, prog.meta_vars['phase']] > 1).all()        # past the fixation phase


class _Kick(object):
    """match_to_sample.py:45-71 in structure: one random speed and sign per call, then a quarter turn of
    every mover's offset from the centre scaled by its distance, given to the mover and to its twin."""

    def __init__(self, speeds):
        self._speeds = speeds

    def reset(self, state, meta_state):
        pass

    def step(self, state, meta_state):
        del meta_state
        w = np.random.uniform(*self._speeds)
        w *= (2 * np.random.randint(2) - 1)
        for a, b in zip(state['movers'], state['twins']):
            rel = a.position - 0.5
            turned = np.matmul(np.array([[0, -1], [1, 0]]), rel)
            v = turned * np.linalg.norm(rel) * w
            a.velocity = v
            b.velocity = v


def test_traced_rule_with_random_draws_and_vector_algebra():
    """np.random.uniform / randint inside a traced rule become rule-noise columns; np.matmul with a
    quarter-turn matrix and np.linalg.norm on a sprite's position are lowered.  With the uniforms behind the
    draws replayed, the oracle's velocities equal what the Python rule computes on the host sprites."""
    import moog_b200  # noqa: F401
    from moog import action_spaces, physics as physics_lib, sprite, tasks
    from moog_b200 import compiler
    from oracle.oracle import Oracle

    def state_initializer():
        movers = [sprite.Sprite(x=0.2 + 0.25 * k, y=0.3 + 0.1 * k, shape='circle', scale=0.05) for k in range(3)]
        twins = [sprite.Sprite(x=m.x, y=m.y, shape='square', scale=0.03) for m in movers]
        return collections.OrderedDict([('movers', movers), ('twins', twins)])

    rule = _Kick((0.1, 0.3))
    cfg = dict(state_initializer=state_initializer, physics=physics_lib.Physics(updates_per_env_step=1),
               task=tasks.CompositeTask(timeout_steps=10), action_space=action_spaces.Grid(action_layers=()),
               observers={}, game_rules=(rule,))
    states = [state_initializer()]
    prog = compiler.compile_config(cfg, states)
    assert prog.rule_noise_dim == 2 and prog.rule_draws[0][1] == [('uniform', 0, None), ('randint', 1, 2)]
    for u0, sign in ((0.25, 0), (0.8, 1)):
        orc = Oracle(prog, compiler.pack_states(prog, states))
        orc.post_reset(rule_noise=np.array([[u0, (sign + 0.5) / 2]]))
        # the same call in Python, its two draws forced to the same outcomes
        twin_state = state_initializer()
        real = (np.random.uniform, np.random.randint)
        np.random.uniform = lambda lo, hi: lo + (hi - lo) * u0
        np.random.randint = lambda n: sign
        try:
            _Kick((0.1, 0.3)).step(twin_state, None)
        finally:
            np.random.uniform, np.random.randint = real
        for layer, lo in (('movers', prog.layer_off[0]), ('twins', prog.layer_off[1])):
            want = np.array([sp.velocity for sp in twin_state[layer]]).T
            assert np.array_equal(orc.dyn[0, 2:4, lo:lo + 3], want), (layer, u0, sign)
        assert (orc.dyn[0, 2:4, :] != 0).any()


@pytest.mark.reference
def test_shipped_match_to_sample_compiles_unchanged(monkeypatch):
    """Build container only: moog_demos/example_configs/match_to_sample.py as shipped, on this repo's `moog`
    package: PhaseSequence of four phases, its own BeginMotion rule (random draws, matmul, norm, zip over two
    layers), metadata reward, a Reset condition on a sprite and meta_state, TetherZippedLayers."""
    import importlib
    import sys
    import moog_b200  # noqa: F401
    from moog_b200 import compiler
    from oracle.oracle import Oracle
    import moog  # noqa: F401  (this repo's package, before the reference's directory joins sys.path)
    monkeypatch.setattr(sys, 'path', sys.path + ['/root/reference'])
    for name in [m for m in sys.modules if m.startswith('moog_demos')]:
        monkeypatch.delitem(sys.modules, name)
    shipped = importlib.import_module('moog_demos.example_configs.match_to_sample')
    np.random.seed(6)
    cfg = shipped.get_config(4)
    states = [cfg['state_initializer']() for _ in range(2)]
    prog = compiler.compile_config(cfg, states)
    assert list(prog.meta_vars) == ['phase'] and prog.meta_keys == ['prey'] and len(prog.rule_draws) == 1
    orc = Oracle(prog, compiler.pack_states(prog, states))
    Oracle.set_seed(9)
    orc.post_reset()
    for _ in range(12):
        orc.step(np.zeros((2, 2)))
    assert (orc.envi[:, 2] == 0).all()
    targets = prog.layer_off[prog.layer_index('targets')]
    assert (np.abs(orc.dyn[:, 2:4, targets:targets + 4]).max(axis=(1, 2)) > 0.01).all(), 'BeginMotion set the targets moving'


def test_modify_meta_state_rules():
    """modify_meta_state.py:8-48: `ModifyMetaState(modifier)` (the modifier traced: it reads and assigns
    entries of the dict) and `UpdateMetaStateValue(key, value)` with a string value."""
    import moog_b200  # noqa: F401
    from moog import game_rules as gr
    from moog_b200 import compiler
    from oracle.oracle import Oracle

    def count_up(meta_state):
        meta_state['count'] += 2
        if meta_state['count'] > 5:
            meta_state['level'] = meta_state['level'] + 1
            meta_state['count'] = 0

    rules = (gr.ModifyMetaState(count_up),
             gr.ConditionalRule(condition=lambda state, meta_state: meta_state['level'] >= 2,
                                rules=gr.UpdateMetaStateValue('phase', 'late')))
    cfg, states = _phase_config(rules, meta=lambda: {'count': 0, 'level': 0, 'phase': 'early'})
    prog = compiler.compile_config(cfg, states)
    orc = Oracle(prog, compiler.pack_states(prog, states))
    orc.post_reset()
    host = {'count': 0, 'level': 0, 'phase': 'early'}
    count_up(host)
    for t in range(9):
        got = {k: orc.envf[0, s] for k, s in prog.meta_vars.items()}
        assert (got['count'], got['level']) == (host['count'], host['level']), t
        assert prog.strings[int(got['phase']) - 1] == host['phase'], t
        orc.step(np.array([[0.5, 0.5]]))
        count_up(host)
        if host['level'] >= 2:
            host['phase'] = 'late'
    assert host['phase'] == 'late' and orc.envi[0, 2] == 0


@pytest.mark.gpu
def test_cuda_phase_sequence_through_the_public_api():
    """The PhaseSequence / Fixation / meta_state config of the CPU test above through BatchedEnvironment
    (`meta_state_initializer` accepted, auto-resets from the pool, Phase durations drawn on the device):
    `env.meta_state(i)` returns the dict the reference would hold; every env walks the phases in order
    and is paid once per episode."""
    import torch
    import moog_b200  # noqa: F401
    from moog import game_rules as gr, tasks
    from moog_b200.batched_env import BatchedEnvironment

    def bump(s):
        s.c0 = s.c0 + 1.

    phases = gr.PhaseSequence(
        gr.Phase(continual_rules=gr.Fixation('agent', 'cross', 0.1, 'held'),
                 end_condition=lambda state, meta_state: meta_state['held'] >= 3, name='fixate'),
        gr.Phase(continual_rules=gr.ModifySprites('agent', bump), duration=lambda: np.random.randint(2, 6), name='count'),
        gr.Phase(name='done'),
        meta_state_phase_name_key='phase')
    task = tasks.CompositeTask(tasks.Reset(condition=lambda state, meta_state: meta_state['phase'] == 'done',
                                           reward_fn=lambda _: 7., steps_after_condition=1), timeout_steps=100)
    cfg, states = _phase_config((phases,), meta=lambda: {'phase': '', 'held': 0}, task=task)
    N = 256
    env = BatchedEnvironment(**cfg, num_envs=N, device='cuda:0', seed=2, initial_states=states)
    env.reset()
    assert env.meta_state(0) == {'phase': 'fixate', 'held': 1}
    act = torch.full((N, 2), 0.5, dtype=torch.float64)
    paid = np.zeros(N)
    lengths = []
    seen = [[] for _ in range(4)]
    for t in range(40):
        ts = env.step(act)
        r = ts.reward.cpu().numpy()
        paid += np.where(np.isnan(r), 0.0, r)
        for e in range(4):
            seen[e].append(env.meta_state(e)['phase'])
    st = env.engine.state.download()
    assert (st['envi'][:, 2] == 0).all()
    episodes = st['envi'][:, 3]
    # (an env whose condition has just held was paid for an episode that ends on the next step)
    assert (episodes >= 3).all() and np.isin(paid - 7.0 * episodes, (0.0, 7.0)).all(), 'paid once per episode'
    assert len(set(episodes.tolist())) > 1, 'the drawn durations differ between envs'
    for e in range(4):
        order = [p for k, p in enumerate(seen[e]) if k == 0 or seen[e][k - 1] != p]
        assert order[:3] == ['fixate', 'count', 'done'], order


_SHIPPED = [('pong', None), ('falling_balls', None), ('colliding_predators', None), ('predators_arena', 3),
            ('cleanup', None), ('chase_avoid_torus', 0), ('pacman', 0), ('parallelogram_catch', 0),
            ('first_person_predators_prey', None), ('red_green', 1), ('bounce_box_contact_prediction', True),
            ('functional_maze', None), ('multi_tracking_with_feature', 3), ('match_to_sample', 4)]


@pytest.mark.reference
@pytest.mark.parametrize('module,level', _SHIPPED)
def test_every_shipped_example_config_compiles_unchanged(monkeypatch, module, level):
    """Build container only: each of the 14 modules of /root/reference/moog_demos/example_configs, imported as
    shipped on this repo's `moog` package, compiles to a valid program and steps on the oracle without an
    error flag.  (`Physics.step` inside two initializers is served by the oracle here -- a TEST stand-in for
    the CUDA call of moog_b200/host_physics.py, which has its own GPU test.)"""
    import ctypes
    import importlib
    import sys
    import moog_b200  # noqa: F401
    from moog_b200 import capi, compiler, host_physics
    from oracle.oracle import Oracle

    def _oracle_step(physics, state):
        from moog import action_spaces, tasks
        cfg = dict(physics=physics, task=tasks.CompositeTask(), action_space=action_spaces.Grid(action_layers=()),
                   observers={}, game_rules=())
        prog = compiler.compile_config(cfg, [state])
        orc = Oracle(prog, compiler.pack_states(prog, [state]))
        orc.physics_step()
        for l, name in enumerate(prog.layer_names):
            for k, sp in enumerate(state[name]):
                s = prog.layer_off[l] + k
                sp.position = np.array(orc.dyn[0, 0:2, s])
                sp.velocity = np.array(orc.dyn[0, 2:4, s])

    monkeypatch.setattr(host_physics, 'step', _oracle_step)
    import moog  # noqa: F401  (this repo's package, before the reference's directory joins sys.path)
    monkeypatch.setattr(sys, 'path', sys.path + ['/root/reference'])
    for name in [m for m in sys.modules if m.startswith('moog_demos')]:
        monkeypatch.delitem(sys.modules, name)
    shipped = importlib.import_module('moog_demos.example_configs.' + module)
    np.random.seed(11)
    cfg = shipped.get_config(level)
    states = [cfg['state_initializer']() for _ in range(2)]
    capacity = {'predators': 24, 'prey': 24} if module == 'first_person_predators_prey' else None
    prog = compiler.compile_config(cfg, states, layer_capacity=capacity)
    buf = (ctypes.c_uint8 * len(prog.blob)).from_buffer_copy(prog.blob)
    assert capi.lib().moog_program_validate(ctypes.cast(buf, ctypes.c_void_p), len(prog.blob)) == 0
    orc = Oracle(prog, compiler.pack_states(prog, states))
    Oracle.set_seed(4)
    orc.post_reset()
    rng = np.random.RandomState(0)
    ad = max(prog.action_dim, 1)
    grid = any(kind == 'Grid' for _, kind, _, _ in prog.action_layout)
    for t in range(12):
        act = rng.randint(0, 5, size=(2, ad)).astype(np.float64) if grid else rng.uniform(0.2, 0.8, size=(2, ad))
        noise = rng.uniform(size=(2, prog.K, prog.noise_dim)) if prog.noise_dim else None
        Oracle.set_seed(100 + t)
        orc.step(act, noise=noise)
    assert (orc.envi[:, 2] == 0).all(), orc.envi[:, 2]


def _small_configs():
    """The hand-made configs of the CPU tests above, as (name, config, states, action width)."""
    import moog_b200  # noqa: F401
    from moog import action_spaces, game_rules as gr, physics as physics_lib, sprite, tasks
    out = []
    # metadata reward with branches
    metadata = [{'goal': g, 'when': w} for g in (0, 1) for w in (5, 31, 60)]
    cfg, states = _config(_answer, metadata)
    out.append(('answer', cfg, states))
    # decision-tree Reset task
    def verdict_state(x, goal):
        boxes = [sprite.Sprite(x=0.3, y=0.5, shape='square', scale=0.1), sprite.Sprite(x=0.7, y=0.5, shape='square', scale=0.1)]
        agent = sprite.Sprite(x=x, y=0.5, shape='circle', scale=0.08, metadata={'goal': goal})
        return collections.OrderedDict([('boxes', boxes), ('agent', [agent])])
    states = [verdict_state(x, goal) for x in (0.3, 0.5, 0.7) for goal in (False, True)]
    cfg = dict(state_initializer=lambda: verdict_state(0.5, True), physics=physics_lib.Physics(updates_per_env_step=1),
               task=tasks.CompositeTask(tasks.Reset(condition=lambda state: _verdict(state) != 0, reward_fn=_verdict,
                                                    steps_after_condition=2), timeout_steps=100),
               action_space=action_spaces.Grid(scaling_factor=0.01, action_layers='agent'), observers={}, game_rules=())
    out.append(('verdict', cfg, states))
    # traced rule class with a loop over a layer
    def tagger_state(n_items):
        agent = sprite.Sprite(x=0.5, y=0.5, shape='square', scale=0.2, mass=1.)
        xs = [0.45, 0.9, 0.55]
        items = [sprite.Sprite(x=xs[k], y=0.5, shape='circle', scale=0.05, c0=float(k)) for k in range(n_items)]
        return collections.OrderedDict([('items', items), ('agent', [agent])])
    states = [tagger_state(n) for n in (3, 0, 1, 2, 3)]
    cfg = dict(state_initializer=lambda: tagger_state(3), physics=physics_lib.Physics(updates_per_env_step=1),
               task=tasks.CompositeTask(timeout_steps=100), action_space=action_spaces.Grid(action_layers=()),
               observers={}, game_rules=(_Tagger(threshold=3),))
    out.append(('tagger', cfg, states))
    # traced rule with random draws and vector algebra
    def kick_state():
        movers = [sprite.Sprite(x=0.2 + 0.25 * k, y=0.3 + 0.1 * k, shape='circle', scale=0.05) for k in range(3)]
        twins = [sprite.Sprite(x=m.x, y=m.y, shape='square', scale=0.03) for m in movers]
        return collections.OrderedDict([('movers', movers), ('twins', twins)])
    cfg = dict(state_initializer=kick_state, physics=physics_lib.Physics(updates_per_env_step=2),
               task=tasks.CompositeTask(timeout_steps=10), action_space=action_spaces.Grid(action_layers=()),
               observers={}, game_rules=(_Kick((0.1, 0.3)),))
    out.append(('kick', cfg, [kick_state() for _ in range(3)]))
    # phases, fixation, meta_state, modify-meta-state rules
    def bump(s):
        s.c0 = s.c0 + 1.

    def count_up(meta_state):
        meta_state['count'] += 2
        if meta_state['count'] > 5:
            meta_state['level'] = meta_state['level'] + 1
            meta_state['count'] = 0
    phases = gr.PhaseSequence(
        gr.Phase(continual_rules=gr.Fixation('agent', 'cross', 0.1, 'held'),
                 end_condition=lambda state, meta_state: meta_state['held'] >= 3, name='fixate'),
        gr.Phase(one_time_rules=gr.UpdateMetaStateValue('level', 10), continual_rules=gr.ModifySprites('agent', bump),
                 duration=lambda: np.random.randint(2, 6), name='count'),
        gr.Phase(name='done'),
        meta_state_phase_name_key='phase')
    task = tasks.CompositeTask(tasks.Reset(condition=lambda state, meta_state: meta_state['phase'] == 'done',
                                           reward_fn=lambda _: 7., steps_after_condition=1), timeout_steps=100)
    cfg, states = _phase_config((phases, gr.ModifyMetaState(count_up)),
                                meta=lambda: {'phase': '', 'held': 0, 'count': 0, 'level': 0}, task=task)
    out.append(('phases', cfg, states * 3))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize('which', ['answer', 'verdict', 'tagger', 'kick', 'phases'])
def test_cuda_matches_oracle_on_the_small_configs(which):
    """The hand-made configs of this file (metadata rewards, decision trees, traced rule classes with loops /
    random draws / vector algebra, PhaseSequence + Fixation + meta_state + ModifyMetaState) on the CUDA path
    against the oracle: same Philox stream, auto-resets from the pool, every array of the record -- envf with
    the rules' own variables, meta_state entries and metadata columns included -- identical after every step."""
    from moog_b200 import compiler
    from moog_b200.batched_env import Engine
    from oracle.oracle import Oracle
    name, cfg, states = [c for c in _small_configs() if c[0] == which][0]
    prog = compiler.compile_config(cfg, states)
    pool_arrays = {k: v for k, v in compiler.pack_states(prog, states).items() if k in util.STATE_KEYS}
    P = len(states)
    N = 8 * P
    rng = np.random.RandomState(3)
    arrays = {k: np.ascontiguousarray(pool_arrays[k][np.arange(N) % P]) for k in util.STATE_KEYS}
    orc, eng = Oracle(prog, arrays), Engine(prog, N, 'cuda:0', seed=17)
    eng.state.upload(arrays)
    eng.set_pool(pool_arrays)
    pool = Oracle(prog, pool_arrays)
    Oracle.set_seed(17)
    orc.post_reset()
    eng.post_reset()
    ad = max(prog.action_dim, 1)
    grid = any(kind == 'Grid' for _, kind, _, _ in prog.action_layout)
    for t in range(25):
        act = rng.randint(0, 5, size=(N, ad)).astype(np.float64) if grid else rng.uniform(0.3, 0.7, size=(N, ad))
        ri = rng.randint(0, P, size=N)
        Oracle.set_seed(eng.call_seed())
        r_ref, st_ref, _ = orc.step_auto(act, pool, ri)
        eng.env_step(act, auto_reset=True, reset_index=ri, want_counters=True)
        dev = eng.state.download()
        assert np.array_equal(eng.step_type.cpu().numpy(), st_ref), (name, t)
        assert _same_f(eng.reward.cpu().numpy(), r_ref.astype(np.float32)), (name, t, 'reward')
        assert np.array_equal(eng.counters.cpu().numpy()[:, :2], orc.counters[:, :2]), (name, t, 'overlap calls')
        assert np.array_equal(dev['cnt'], orc.cnt) and np.array_equal(dev['envi'][:, :6], orc.envi[:, :6]), (name, t)
        assert np.array_equal(dev['envf'], orc.envf, equal_nan=True), (name, t, 'envf')
        for e in range(N):
            util.assert_live_equal(prog, {k: dev[k][e] for k in ('dyn', 'stat', 'meta', 'vtx', 'cnt')},
                                   {k: getattr(orc, k)[e] for k in ('dyn', 'stat', 'meta', 'vtx', 'cnt')},
                                   '{} step {} env {}'.format(name, t, e))
    assert (orc.envi[:, 2] == 0).all()


def _same_f(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.nan_to_num(a), np.nan_to_num(b))
