"""Golden vectors for `game_rules.CreateSprites` behind `np.random.binomial` conditions: the UNMODIFIED
reference stepped through scenes whose sprites appear during the episode
(game_rules/create_sprites.py:8-34, conditional.py:55-58, state_initialization/sprite_generators.py:26-105).

TEST INFRASTRUCTURE ONLY; runs in the build container only (needs /root/reference, imported
through oracle/shims).  Usage:

    python oracle/gen_golden_spawn.py [scene ...]    # writes tests/golden/spawn_*.npz

The reference draws from np.random's global stream; the batched implementation keys its draws by
(seed, env, counters).  What is recorded here lets the oracle REPLAY the reference's draws
(oracle.Oracle.force_factors) so that everything around the draw is pinned bit for bit:

    blob, layer_names, init_* / reset_*    as gen_golden.py
    rule_noise[T + 1, C]    0.0 where the condition `np.random.binomial(1, p)` of column c returned 1 in
                            the rules pass of step t, 1.0 where it returned 0; row 0 is the pass
                            Environment.reset() makes (environment.py:94-95), row t + 1 step t
    factors[M, 14], factor_start[T + 2]     the dict every factor_dist.sample() call returned, in call
                            order (x y x_vel y_vel angle angle_vel mass scale aspect_ratio c0 c1 c2
                            opacity, then the shape's index in the program's shape records); the calls of
                            pass p are rows factor_start[p] : factor_start[p + 1]
    actions[T, A], reward[T], last[T], dyn[T], cnt[T]
    n_calls[T + 1], n_true[T + 1], true_hash[T + 1]     Sprite.overlaps_sprite calls of every pass (row 0:
                            the reset pass), slots resolved at call time -- a candidate that is not in the
                            state yet counts as the slot it would take
    stat / meta / vtx at `full_steps[F]`;  frames[G] at `frame_steps[G]` (-1: after reset)
"""

import importlib
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:] = [p for p in sys.path if os.path.abspath(p or '.') != os.path.join(_ROOT, 'oracle')]
sys.path.insert(0, _ROOT)

from oracle import refenv  # noqa: E402

refenv.activate()

import moog_b200  # noqa: E402,F401
from moog_b200 import compiler  # noqa: E402
from moog import environment  # noqa: E402
from moog import sprite as sprite_lib  # noqa: E402
from oracle.gen_golden import true_event_hash, _flat_action  # noqa: E402

SCENES = {
    # name: (module, seed, T, layer_capacity or None -> the module's LAYER_CAPACITY)
    'zoo': ('moog_b200.configs.spawn_zoo', 31, 59, None),
    'first_person': ('moog_demos.example_configs.first_person_predators_prey', 32, 110,
                     {'predators': 44, 'prey': 28}),
}


def _create_rules(rules):
    out = []
    for r in rules:
        if type(r).__name__ == 'CreateSprites':
            out.append(r)
        elif hasattr(r, '_rules'):
            out += _create_rules(r._rules)  # pylint: disable=protected-access
    return out


def generate(name, out_dir):
    module, seed, T, capacity = SCENES[name]
    np.random.seed(seed)
    mod = importlib.import_module(module)
    config = mod.get_config(None)
    capacity = capacity if capacity is not None else mod.LAYER_CAPACITY
    env = environment.Environment(**config)
    init_state = env.state_initializer()
    prog = compiler.compile_config(config, [init_state], layer_capacity=capacity)
    init = compiler.pack_states(prog, [init_state])
    table = init['shape_table']
    env.state_initializer = lambda: init_state
    renderer = config.get('observers', {}).get('image')
    attr_keys = compiler._ATTR_KEYS  # pylint: disable=protected-access
    defaults = compiler._sprite_defaults()  # pylint: disable=protected-access

    # -- instrumentation ---------------------------------------------------------------------
    bern, rows = [], []
    ctx = {'rule': None, 'base': 0, 'made': 0}
    orig_binomial = np.random.binomial

    def _binomial(n, p, size=None):
        r = orig_binomial(n, p, size)
        bern.append(int(r))
        return r

    def _shape_id(shape):
        key = shape if isinstance(shape, str) else np.asarray(shape, dtype=np.float64).tobytes()
        return prog.z_shape_ids[key]

    creates = _create_rules(config['game_rules'])
    for rule in creates:
        dist = compiler._Recipe(rule._generator).factor_dist  # pylint: disable=protected-access

        def _sample(rng=None, _orig=dist.sample):
            out = _orig(rng)
            row = [float(out.get(k, defaults[k])) for k in attr_keys] + [float(_shape_id(out.get('shape', defaults['shape'])))]
            for k, v in zip(attr_keys, row):          # the record keeps doubles: every factor must survive that
                assert v == out.get(k, defaults[k]), (k, out.get(k))
            rows.append(row)
            return out
        dist.sample = _sample

        def _step(state, meta_state, _rule=rule, _orig=rule.step):
            ctx['rule'], ctx['base'] = _rule, len(state[_rule._layer])  # pylint: disable=protected-access
            try:
                return _orig(state, meta_state)
            finally:
                ctx['rule'] = None
        rule.step = _step

    calls = []
    orig_overlaps = sprite_lib.Sprite.overlaps_sprite

    def _slot_now(sp):
        for l, lname in enumerate(prog.layer_names):
            for k, other in enumerate(env.state[lname]):
                if other is sp:
                    return int(prog.layer_off[l]) + k
        rule = ctx['rule']
        assert rule is not None, 'an overlap call on a sprite outside the state, outside CreateSprites'
        assert compiler._Recipe(rule._generator).num_sprites == 1, 'candidate slots are only tracked for one sprite per call'  # pylint: disable=protected-access
        return int(prog.layer_off[prog.layer_index(rule._layer)]) + ctx['base']  # pylint: disable=protected-access

    def _overlaps(this, other):
        r = bool(orig_overlaps(this, other))
        calls.append((_slot_now(this), _slot_now(other), r))
        return r

    def _instrumented(fn):
        del bern[:], calls[:]
        n_rows = len(rows)
        np.random.binomial = _binomial
        sprite_lib.Sprite.overlaps_sprite = _overlaps
        try:
            out = fn()
        finally:
            np.random.binomial = orig_binomial
            sprite_lib.Sprite.overlaps_sprite = orig_overlaps
        assert len(bern) == prog.rule_noise_dim, 'one binomial draw per Bernoulli condition and pass expected'
        h, n_true = 0, 0
        for a, b, r in calls:
            if r:
                n_true += 1
                h = true_event_hash(h, a, b)
        return out, [0.0 if b else 1.0 for b in bern], len(rows) - n_rows, (len(calls), n_true, np.uint64(h))

    # -- trajectory ---------------------------------------------------------------------------
    rec = {k: [] for k in ('rule_noise', 'n_calls', 'n_true', 'true_hash', 'actions', 'reward', 'last', 'dyn', 'cnt')}
    full = {k: [] for k in ('stat', 'meta', 'vtx')}
    full_steps, frames, frame_steps, factor_start = [], [], [], [0]

    def _note_pass(rn, counters):
        rec['rule_noise'].append(rn)
        for k, v in zip(('n_calls', 'n_true', 'true_hash'), counters):
            rec[k].append(v)
        factor_start.append(len(rows))

    ts, rn, _, counters = _instrumented(env.reset)
    _note_pass(rn, counters)
    after_reset = compiler.pack_states(prog, [env.state], table)
    if renderer is not None:
        frames.append(np.asarray(ts.observation['image']))
        frame_steps.append(-1)
    for t in range(T):
        action = env.action_space.random_action()
        flat = _flat_action(prog, action)
        ts, rn, _, counters = _instrumented(lambda: env.step(action))
        _note_pass(rn, counters)
        st = compiler.pack_states(prog, [env.state], table)
        rec['actions'].append(flat)
        rec['reward'].append(0.0 if ts.reward is None else float(ts.reward))
        rec['last'].append(bool(ts.last()))
        rec['dyn'].append(st['dyn'][0])
        rec['cnt'].append(st['cnt'][0])
        if t % 4 == 0 or t == T - 1 or ts.last():
            full_steps.append(t)
            for k in full:
                full[k].append(st[k][0])
        if renderer is not None and (t % 10 == 0 or t == T - 1 or ts.last()):
            frames.append(np.asarray(ts.observation['image']))
            frame_steps.append(t)
        if ts.last():
            break
    assert max(rec['cnt'][-1]) > 3 and len(rows) > 10, 'nothing was created'
    caps = np.array(prog.layer_cap)
    assert all((c[:len(caps)] <= caps).all() for c in rec['cnt']), 'a layer outgrew its capacity: raise it'

    out = dict(blob=np.frombuffer(prog.blob, dtype=np.uint8), layer_names=np.array(prog.layer_names),
               factors=np.array(rows, dtype=np.float64).reshape(-1, 14), factor_start=np.array(factor_start, dtype=np.int32),
               full_steps=np.array(full_steps, dtype=np.int32),
               frames=np.array(frames, dtype=np.uint8), frame_steps=np.array(frame_steps, dtype=np.int32))
    for k in ('dyn', 'stat', 'meta', 'vtx', 'cnt', 'envi', 'envf'):
        out['init_' + k] = init[k][0]
        out['reset_' + k] = after_reset[k][0]
    for k, v in rec.items():
        out[k] = np.array(v)
    for k, v in full.items():
        out[k] = np.array(v)
    path = os.path.join(out_dir, 'spawn_' + name + '.npz')
    np.savez_compressed(path, **out)
    print('{:14s} T={:3d} slots={:3d} sample() calls={} final counts={} true/pass={:.1f} -> {} ({} KB)'.format(
        name, len(rec['reward']), prog.n_slots, len(rows), rec['cnt'][-1][:prog.n_layers].tolist(),
        float(np.mean(rec['n_true'])), path, os.path.getsize(path) // 1024))


RESET_SCENES = {
    # name: (module, seed, number of initial states)
    'colliding_predators84': ('moog_b200.configs.colliding_predators84', 33, 12),
}


def shape_key(shape):
    """Text key of a shape factor: 's:<name>' or 'a:<hex of the float64 vertex array>'."""
    return 's:' + shape if isinstance(shape, str) else 'a:' + np.asarray(shape, dtype=np.float64).tobytes().hex()


def generate_resets(name, out_dir):
    """The reference's state initializer (sprite_generators.py:75-103 with disjoint / without_overlapping)
    called P times; every OUTERMOST factor_dist.sample() call is recorded, so that the oracle's reset
    sampler can replay the draws:  factors[M, 13], shape_keys[M], row_start[P + 1], and the packed states
    pool_*[P, ...] the initializer returned (layout words layer_off / voff stored for the test to check
    that its own program -- compiled with the reset sampler, outside this script -- agrees)."""
    from moog.state_initialization import distributions as distribs
    module, seed, P = RESET_SCENES[name]
    np.random.seed(seed)
    config = importlib.import_module(module).get_config(None)
    attr_keys = compiler._ATTR_KEYS  # pylint: disable=protected-access
    defaults = compiler._sprite_defaults()  # pylint: disable=protected-access
    rows, keys, depth = [], [], [0]
    patched = []
    for cls_name in ('Continuous', 'Discrete', 'Mixture', 'Intersection', 'Product', 'SetMinus', 'Selection',
                     'DependentDistribution'):
        cls = getattr(distribs, cls_name)

        def _sample(self, rng=None, _orig=cls.sample):
            depth[0] += 1
            try:
                out = _orig(self, rng)
            finally:
                depth[0] -= 1
            if depth[0] == 0:
                row = [float(out.get(k, defaults[k])) for k in attr_keys]
                for k, v in zip(attr_keys, row):
                    assert v == out.get(k, defaults[k]), (k, out.get(k))
                rows.append(row)
                keys.append(shape_key(out.get('shape', defaults['shape'])))
            return out
        patched.append((cls, cls.sample))
        cls.sample = _sample
    row_start, states = [0], []
    try:
        for _ in range(P):
            states.append(config['state_initializer']())
            row_start.append(len(rows))
    finally:
        for cls, orig in patched:
            cls.sample = orig
    prog = compiler.compile_config(config, states)
    pool = compiler.pack_states(prog, states)
    out = dict(factors=np.array(rows, dtype=np.float64), shape_keys=np.array(keys), row_start=np.array(row_start, dtype=np.int32),
               layer_off=np.array(prog.layer_off, dtype=np.int32), voff=np.array(prog.voff, dtype=np.int32),
               layer_names=np.array(prog.layer_names))
    for k in ('dyn', 'stat', 'meta', 'vtx', 'cnt'):
        out['pool_' + k] = pool[k]
    path = os.path.join(out_dir, 'resets_' + name + '.npz')
    np.savez_compressed(path, **out)
    print('{:22s} {} initial states, {} sample() calls -> {} ({} KB)'.format(name, P, len(rows), path,
                                                                             os.path.getsize(path) // 1024))


def main():
    out_dir = os.path.join(_ROOT, 'tests', 'golden')
    for n in (sys.argv[1:] or list(SCENES) + list(RESET_SCENES)):
        if n in RESET_SCENES:
            generate_resets(n, out_dir)
        else:
            generate(n, out_dir)


if __name__ == '__main__':
    main()
