"""`game_rules.CreateSprites` behind `np.random.binomial(1, p)` conditions
(/root/reference/moog/game_rules/create_sprites.py:8-34, conditional.py:55-58; the pattern of
moog_demos/example_configs/first_person_predators_prey.py:176-193).

The reference draws from np.random's global stream in Python call order; a batch of envs stepped
concurrently keys every draw by (seed, env, counters) instead (Philox).  So:
  * the LOGIC (which sprites a new one must avoid, where it is appended, fail_gracefully, what a
    rejected try costs, Mixture / SetMinus factor trees, the rules that follow in the same pass) is
    pinned against the unmodified reference on a trajectory whose draws are replayed into the
    oracle (tests/golden/spawn_zoo.npz, first_person.npz: test_oracle_replays_reference_*);
  * the CUDA path is compared with the oracle draw for draw on the Philox stream (gpu tests);
  * that the Philox draws follow the reference's DISTRIBUTIONS is checked on the factors of the
    sprites made (test_cuda_created_factors_follow_the_distributions).
"""
import importlib

import numpy as np
import pytest

from tests import util


def _compile(name='spawn_zoo', n_states=6, seed=3):
    import moog_b200  # noqa: F401
    from moog_b200 import compiler
    mod = importlib.import_module('moog_b200.configs.' + name)
    cfg = mod.get_config()
    np.random.seed(seed)
    states = [cfg['state_initializer']() for _ in range(n_states)]
    prog = compiler.compile_config(cfg, states, layer_capacity=mod.LAYER_CAPACITY)
    pool = compiler.pack_states(prog, states)
    return cfg, prog, {k: pool[k] for k in util.STATE_KEYS}


def test_create_sprites_lowering():
    """The program holds one MOOG_R_CREATE_SPRITES op per rule, guarded by a MOOG_SC_BERNOULLI condition
    with a rule-noise column of its own; the growing layers get the room and outline width asked for."""
    from moog_b200 import compiler as C
    _, prog, _ = _compile()
    kinds = [o['kind'] for o in prog.ops]
    assert kinds.count(C.R_CREATE_SPRITES) == 2 and kinds.count(C.SC_BERNOULLI) == 2
    bern = [o for o in prog.ops if o['kind'] == C.SC_BERNOULLI]
    assert sorted(o['i'][0] for o in bern) == [0, 1] and sorted(o['p'][0] for o in bern) == [0.35, 0.6]
    assert prog.rule_noise_dim == 2
    assert prog.layer_cap == [2, 32, 32, 1] and prog.layer_vcap[1] == 30 and prog.layer_vcap[2] == 4
    drops = [o for o in prog.ops if o['kind'] == C.R_CREATE_SPRITES and o['i'][0] == 1][0]
    assert drops['flags'] & C.FL_FAIL_GRACEFULLY and drops['i'][1] == 1 and drops['p'][0] == 6.0
    assert [prog.ipool[drops['i'][2] + q] for q in range(drops['i'][3])] == [0, 3, 1]      # walls, agent, drops


def test_create_sprites_is_refused_where_it_cannot_be_lowered():
    import moog_b200  # noqa: F401
    from moog import game_rules
    from moog_b200 import compiler
    from moog_b200.configs import spawn_zoo
    cfg = spawn_zoo.get_config()
    states = [cfg['state_initializer']()]
    bad = dict(cfg, game_rules=(game_rules.CreateSprites('drops', lambda without_overlapping: []),))
    with pytest.raises(compiler.CompileError, match='generate_sprites'):
        compiler.compile_config(bad, states)
    bad = dict(cfg, game_rules=(game_rules.ConditionalRule(
        condition=lambda state: np.random.binomial(2, 0.5), rules=cfg['game_rules'][2]),))
    with pytest.raises(Exception):
        compiler.compile_config(bad, states)


def test_oracle_create_sprites_invariants():
    """On the oracle alone (Philox draws): new sprites never overlap what they must avoid at the moment
    they appear, are appended in order, the Bernoulli conditions fire at their rates, a full layer
    flags MOOG_ERR_LAYER_OVERFLOW and creates nothing."""
    from oracle.oracle import Oracle
    _, prog, pool = _compile()
    N = 48
    rng = np.random.RandomState(2)
    arrays = {k: np.ascontiguousarray(pool[k][rng.randint(0, 6, size=N)]) for k in util.STATE_KEYS}
    orc = Oracle(prog, arrays)
    Oracle.set_seed(11)
    orc.post_reset()
    lo = prog.layer_off
    fired = np.zeros(2)
    passes = 0
    for t in range(50):
        before = orc.cnt.copy()
        created = orc.envi[:, 4].copy()
        Oracle.set_seed(1000 + t)
        orc.step(rng.uniform(-1, 1, size=(N, 2)))
        passes += N
        # drops: at most one more, minus those the agent ate in this pass (VanishOnContact runs after)
        assert (orc.cnt[:, 1] <= before[:, 1] + 1).all()
        fired[0] += (orc.envi[:, 4] - created >= 1).sum()
        ov = orc.overlap_pairs('drops', 'walls')
        assert not ov.any(), 'a drop appeared on a wall'
        dd = orc.overlap_pairs('drops', 'drops')
        for e in range(N):
            n = orc.cnt[e, 1]
            assert not (dd[e][:n, :n] & ~np.eye(n, dtype=bool)).any(), 'two drops overlap'
        # sparks are axis-aligned squares on the border with speeds in the ring
        for e in range(N):
            n = orc.cnt[e, 2]
            v = orc.dyn[e][2:4, lo[2]:lo[2] + n]
            assert (np.abs(v).max(axis=0) >= 0.02).all() and (np.abs(v) < 0.05).all()
    assert (orc.envi[:, 2] == 0).all()
    # tiny layer: the flag is raised, nothing is written past the layer
    import moog_b200  # noqa: F401
    from moog_b200 import compiler
    from moog_b200.configs import spawn_zoo
    cfg = spawn_zoo.get_config()
    np.random.seed(3)
    states = [cfg['state_initializer']() for _ in range(2)]
    small = compiler.compile_config(cfg, states, layer_capacity={'drops': 3, 'sparks': 2})
    orc = Oracle(small, compiler.pack_states(small, states))
    Oracle.set_seed(1)
    orc.post_reset()
    for t in range(30):
        Oracle.set_seed(50 + t)
        orc.step(np.zeros((2, 2)))
    assert (orc.cnt[:, 1] <= 3).all() and (orc.cnt[:, 2] <= 2).all()
    assert ((orc.envi[:, 2] & 8) != 0).all()


def _same(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.nan_to_num(a), np.nan_to_num(b))


@pytest.mark.gpu
def test_cuda_create_sprites_matches_oracle():
    """spawn_zoo, 192 envs, 140 steps through two time-outs with auto-resets from a pool: the CUDA path
    and the oracle draw from the same Philox stream (io.seed), so the states -- every created sprite's
    factors, outline, slot -- the rewards, step types, overlap pair sets, error words and the frames
    must be IDENTICAL every step (no sin / cos on this path: bit-exact)."""
    from moog_b200.batched_env import Engine
    from oracle.oracle import Oracle
    _, prog, pool_arrays = _compile()
    N = 192
    rng = np.random.RandomState(4)
    arrays = {k: np.ascontiguousarray(pool_arrays[k][rng.randint(0, 6, size=N)]) for k in util.STATE_KEYS}
    orc = Oracle(prog, arrays)
    eng = Engine(prog, N, 'cuda:0', seed=21)
    eng.state.upload(arrays)
    eng.set_pool(pool_arrays)
    pool = Oracle(prog, pool_arrays)
    Oracle.set_seed(21 & 0x7fffffff)            # moog_env_post_reset: the program's "seed" option
    orc.post_reset()
    eng.post_reset()
    orc.envi[:, 0] = rng.randint(0, 40, size=N)        # staggered time-outs
    eng.state.upload(orc.arrays())
    made = 0
    for t in range(140):
        ri = rng.randint(0, 6, size=N)
        act = rng.uniform(-1, 1, size=(N, 2))
        serial = orc.envi[:, 4].copy()
        Oracle.set_seed(eng.call_seed())
        r_ref, st_ref, d_ref = orc.step_auto(act, pool, ri)
        eng.env_step(act, auto_reset=True, reset_index=ri, want_counters=True, frames=t % 20 == 0 or None)
        made += int((orc.envi[:, 4] - serial).sum())
        assert np.array_equal(eng.step_type.cpu().numpy(), st_ref), t
        assert _same(eng.reward.cpu().numpy(), r_ref.astype(np.float32)), (t, 'reward')
        assert np.array_equal(eng.counters.cpu().numpy()[:, :4], orc.counters), (t, 'overlap pair sets')
        dev = eng.state.download()
        assert np.array_equal(dev['cnt'], orc.cnt), t
        assert np.array_equal(dev['envi'][:, :6], orc.envi[:, :6]), (t, 'counters / error words')
        for e in range(N):
            util.assert_live_equal(prog, {k: dev[k][e] for k in ('dyn', 'stat', 'meta', 'vtx', 'cnt')},
                                   {k: getattr(orc, k)[e] for k in ('dyn', 'stat', 'meta', 'vtx', 'cnt')},
                                   'step {} env {}'.format(t, e))
        if t % 20 == 0:
            assert np.array_equal(eng.frames.cpu().numpy(), orc.render()), (t, 'frames')
    assert made > 100 * N // 2, 'CreateSprites calls made: {}'.format(made)
    assert (orc.cnt[:, 1] > 3).any() and (orc.envi[:, 3] >= 2).all()


SPAWN_SCENES = ['zoo', 'first_person']


def _load_spawn(name):
    g = dict(np.load(util.GOLDEN + '/spawn_' + name + '.npz'))
    g['program'] = util.ProgramStub(g['blob'], g['layer_names'])
    return g


def _meta_rows(meta, live):
    """dtype flags and outline sizes of the live sprites.  Row 0 (the shape's number) is left out: a
    created sprite carries its index in the sampler's shape records, a packed reference state the
    index in pack_states' own table -- read only by the scale / aspect_ratio setters, which the
    compiler refuses to combine with the device sampler."""
    return util.canonical_meta(meta, live)[1:]


@pytest.mark.parametrize('name', SPAWN_SCENES)
def test_oracle_replays_reference_create_sprites(name):
    """The unmodified reference's trajectory (oracle/gen_golden_spawn.py) replayed on the oracle: the
    Bernoulli conditions receive the reference's outcomes through the rule-noise columns and every
    generate_sprites try the factor dict the reference's factor_dist.sample() returned.  State (every
    created sprite's attributes, dtype flags, outline, slot), rewards, terminations, the overlap calls
    of every pass (count, Trues, order-sensitive hash -- candidates that were rejected included) and the
    frames are identical, and every recorded draw is consumed by exactly the pass that made it."""
    from oracle.oracle import Oracle
    g = _load_spawn(name)
    prog = g['program']
    try:
        orc = Oracle(prog, util.state_at(g, 0, prefix='init'))
        p = 0
        Oracle.force_factors(g['factors'][g['factor_start'][p]:g['factor_start'][p + 1]])
        orc.post_reset(rule_noise=g['rule_noise'][p][None])
        assert Oracle.forced_left() == 0
        want = util.state_at(g, 0, prefix='reset')
        live = util.live_mask(prog, want['cnt'][0])
        assert np.array_equal(orc.cnt[0], want['cnt'][0])
        assert np.array_equal(orc.dyn[0][:, live], want['dyn'][0][:, live])
        assert np.array_equal(_meta_rows(orc.meta[0], live), _meta_rows(want['meta'][0], live))
        fi = {int(t): k for k, t in enumerate(g['frame_steps'])}
        fu = {int(t): k for k, t in enumerate(g['full_steps'])}
        assert np.array_equal(orc.render()[0], g['frames'][fi[-1]])
        created = 0
        for t in range(len(g['reward'])):
            p = t + 1
            rows = g['factors'][g['factor_start'][p]:g['factor_start'][p + 1]]
            Oracle.force_factors(rows)
            created += len(rows)
            reward, step_type = orc.step(g['actions'][t][None], rule_noise=g['rule_noise'][p][None])
            assert Oracle.forced_left() == 0, (t, 'draws left over')
            assert reward[0] == g['reward'][t] and bool(step_type[0] == 2) == bool(g['last'][t]), t
            assert np.array_equal(orc.cnt[0], g['cnt'][t]), (t, orc.cnt[0], g['cnt'][t])
            live = util.live_mask(prog, g['cnt'][t])
            assert np.array_equal(orc.dyn[0][:, live], g['dyn'][t][:, live]), t
            assert tuple(int(v) for v in orc.counters[0][[0, 1]]) == (int(g['n_calls'][p]), int(g['n_true'][p])), (t, 'overlap calls')
            assert np.uint64(orc.counters[0][3]) == g['true_hash'][p], (t, 'order of the True overlap events')
            if t in fu:
                k = fu[t]
                assert np.array_equal(orc.stat[0][:, live], g['stat'][k][:, live]), t
                assert np.array_equal(_meta_rows(orc.meta[0], live), _meta_rows(g['meta'][k], live)), t
                vlive = util.live_vertex_mask(prog, g['cnt'][t], g['meta'][k])
                assert np.array_equal(orc.vtx[0][vlive], g['vtx'][k][vlive]), t
            if t in fi:
                assert np.array_equal(orc.render()[0], g['frames'][fi[t]]), (t, 'frame')
        assert created > 50 and (orc.envi[0, 2] == 0)
    finally:
        Oracle.force_factors(None)


@pytest.mark.gpu
def test_cuda_first_person_predators_prey_matches_oracle():
    """The SHIPPED first_person_predators_prey config (its program compiled from the reference's own
    config objects travels in the golden): 96 envs x 150 steps, CUDA vs oracle on the same Philox
    stream -- predators and prey appearing on the border (Mixture + SetMinus factors), vanishing far
    away, KeepNearCenter, the `reward_fn` rewards, and the FirstPersonAgent frames -- identical."""
    from moog_b200.batched_env import Engine
    from oracle.oracle import Oracle
    g = _load_spawn('first_person')
    prog = g['program']
    N = 96
    arrays = util.tile_state(util.state_at(g, 0, prefix='init'), N)
    orc = Oracle(prog, arrays)
    eng = Engine(prog, N, 'cuda:0', seed=8)
    eng.state.upload(arrays)
    eng.set_pool({k: arrays[k][:2] for k in util.STATE_KEYS})
    pool = Oracle(prog, {k: arrays[k][:2] for k in util.STATE_KEYS})
    Oracle.set_seed(8)
    orc.post_reset()
    eng.post_reset()
    rng = np.random.RandomState(6)
    ri = np.zeros(N, dtype=np.int32)
    rewards = 0
    for t in range(150):
        act = rng.uniform(-1, 1, size=(N, 2))
        Oracle.set_seed(eng.call_seed())
        r_ref, st_ref, _ = orc.step_auto(act, pool, ri)
        want_frame = t % 25 == 24
        eng.env_step(act, auto_reset=True, reset_index=ri, want_counters=True, frames=want_frame or None)
        assert np.array_equal(eng.step_type.cpu().numpy(), st_ref), t
        assert _same(eng.reward.cpu().numpy(), r_ref.astype(np.float32)), (t, 'reward')
        rewards += int((np.nan_to_num(r_ref) != 0).sum())
        assert np.array_equal(eng.counters.cpu().numpy()[:, :4], orc.counters), (t, 'overlap pair sets')
        dev = eng.state.download()
        assert np.array_equal(dev['cnt'], orc.cnt), t
        assert np.array_equal(dev['envi'][:, :6], orc.envi[:, :6]), t
        for e in range(0, N, 5):
            util.assert_live_equal(prog, {k: dev[k][e] for k in ('dyn', 'stat', 'meta', 'vtx', 'cnt')},
                                   {k: getattr(orc, k)[e] for k in ('dyn', 'stat', 'meta', 'vtx', 'cnt')},
                                   'step {} env {}'.format(t, e))
        if want_frame:
            assert np.array_equal(eng.frames.cpu().numpy(), orc.render()), (t, 'frames')
    # (the fixture's 44 predator slots fill up in some envs late in the run: MOOG_ERR_LAYER_OVERFLOW, on both sides)
    assert ((orc.envi[:, 2] & ~8) == 0).all() and rewards > 20
    assert orc.cnt[:, 3].max() > 30 and orc.cnt[:, 1].max() > 10


@pytest.mark.gpu
def test_cuda_created_factors_follow_the_distributions():
    """The Philox draws against the reference's distributions (spawn_zoo, 2048 envs x 40 steps, the
    BatchedEnvironment API): Bernoulli rates of the two conditions (conditional.py:55-58 with
    np.random.binomial(1, p)), the Mixture's side probabilities and the SetMinus ring of the sparks
    (distributions.py:159-209, 319-365), uniform Discrete shapes and float32 Continuous scales of the
    first drop of an episode (:78-157) -- each within 5 sigma of its expectation."""
    import torch
    import moog_b200  # noqa: F401
    from moog_b200.batched_env import BatchedEnvironment
    from moog_b200.configs import spawn_zoo
    cfg = spawn_zoo.get_config()
    N, T = 2048, 40
    env = BatchedEnvironment(**cfg, num_envs=N, device='cuda:0', seed=3, pool_size=8,
                             layer_capacity=spawn_zoo.LAYER_CAPACITY)
    env.reset()
    prog, eng = env.program, env.engine
    lo = prog.layer_off
    sides = np.zeros(4)
    speeds_ok = True
    n_sparks = 0
    shapes = np.zeros(4)
    scales = []
    drop_calls = spark_calls = 0
    prev = eng.state.download()
    for t in range(T):
        env.step(torch.zeros((N, 2), dtype=torch.float64))
        st = eng.state.download()
        fresh = st['envi'][:, 0] > 0                 # envs that did not auto-reset in this step
        # sparks are appended two at a time and only leave from the far outside: new ones sit at the end
        d_sp = st['cnt'][:, 2] - prev['cnt'][:, 2]
        for e in np.nonzero(fresh & (d_sp == 2))[0][:400]:
            k = lo[2] + st['cnt'][e, 2] - 2
            for s in (k, k + 1):
                # position after one step of flight: undo it
                x = st['dyn'][e, 0, s] - st['dyn'][e, 2, s]
                y = st['dyn'][e, 1, s] - st['dyn'][e, 3, s]
                side = [abs(x + 0.1) < 1e-6, abs(x - 1.1) < 1e-6, abs(y + 0.1) < 1e-6, abs(y - 1.1) < 1e-6]
                assert sum(side) >= 1, (x, y)
                sides[int(np.argmax(side))] += 1
                v = np.abs(st['dyn'][e, 2:4, s])
                speeds_ok &= bool(v.max() >= 0.02 and v.max() < 0.05)
                n_sparks += 1
        spark_calls += int((fresh & (d_sp == 2)).sum())
        d_dr = st['cnt'][:, 1] - prev['cnt'][:, 1]
        # (only the FIRST drop of an episode: later ones are what the rejection against the other drops
        # left over, which favours small shapes)
        for e in np.nonzero(fresh & (d_dr == 1) & (prev['cnt'][:, 1] == 0))[0]:
            s = lo[1] + st['cnt'][e, 1] - 1
            shapes[st['meta'][e, 0, s]] += 1
            scales.append(st['stat'][e, 1, s])
        prev = st
    assert (st['envi'][:, 2] == 0).all()
    assert speeds_ok and n_sparks > 2000
    p = np.array([0.1, 0.2, 0.3, 0.4])
    sigma = np.sqrt(n_sparks * p * (1 - p))
    assert (np.abs(sides - n_sparks * p) < 5 * sigma).all(), (sides, n_sparks)
    n_drops = shapes.sum()
    assert n_drops > 500 and (np.abs(shapes - n_drops / 4) < 5 * np.sqrt(n_drops * 0.1875)).all(), shapes
    scales = np.array(scales)
    assert np.array_equal(scales, scales.astype(np.float32).astype(np.float64))
    assert scales.min() >= 0.1 and scales.max() < 0.17
    assert abs(scales.mean() - 0.135) < 5 * (0.07 / np.sqrt(12)) / np.sqrt(len(scales))
    # Bernoulli rates over all rule passes: serial of CreateSprites calls per env / passes
    calls = st['envi'][:, 4].sum()
    passes = st['envi'][:, 5].sum()
    rate = calls / passes
    assert abs(rate - 0.95) < 5 * np.sqrt((0.6 * 0.4 + 0.35 * 0.65) / passes), (rate, passes)


def test_oracle_replays_reference_state_initializer():
    """The reset sampler's logic against the unmodified reference: colliding_predators' initializer
    (5 predators `disjoint=True, without_overlapping=walls`, then the agent avoiding walls + predators;
    sprite_generators.py:75-103) called 12 times by the reference with every factor_dist.sample()
    recorded (64 of the 136 draws were rejected ones); the oracle's MOOG_Z_GENERATE groups replay those
    draws and must end with the very states the reference returned -- positions, rotated outlines,
    circumscribed radii, inertias, dtype flags -- and consume exactly the draws of each call."""
    import moog_b200  # noqa: F401
    from moog_b200 import compiler
    from moog_b200.configs import colliding_predators84
    from oracle.oracle import Oracle
    g = dict(np.load(util.GOLDEN + '/resets_colliding_predators84.npz'))
    cfg = colliding_predators84.get_config()
    np.random.seed(1)
    states = [cfg['state_initializer']() for _ in range(16)]
    prog = compiler.compile_config(cfg, states, reset_sampler=True)
    assert np.array_equal(g['layer_off'], prog.layer_off) and np.array_equal(g['voff'], prog.voff), 'layouts differ'
    template = compiler.pack_states(prog, [prog.reset_template])
    template = {k: template[k] for k in util.STATE_KEYS}

    def _shape_id(key):
        key = str(key)
        return prog.z_shape_ids[key[2:] if key.startswith('s:') else bytes.fromhex(key[2:])]

    ids = np.array([_shape_id(k) for k in g['shape_keys']], dtype=np.float64)
    rows = np.concatenate([g['factors'], ids[:, None]], axis=1)
    pool = Oracle(prog, template)
    Oracle.set_sample_resets(True)
    try:
        for p in range(len(g['row_start']) - 1):
            arrays = {k: v.copy() for k, v in template.items()}
            arrays['envi'][:, 1] = 1
            orc = Oracle(prog, arrays)
            Oracle.force_factors(rows[g['row_start'][p]:g['row_start'][p + 1]])
            _, step_type, _ = orc.step_auto(None, pool, np.zeros(1, dtype=np.int32))
            assert Oracle.forced_left() == 0 and step_type[0] == 0 and orc.envi[0, 2] == 0, p
            want = {k: g['pool_' + k][p] for k in ('dyn', 'stat', 'meta', 'vtx', 'cnt')}
            assert np.array_equal(orc.cnt[0], want['cnt']), p
            live = util.live_mask(prog, want['cnt'])
            assert np.array_equal(orc.dyn[0][:, live], want['dyn'][:, live]), p
            assert np.array_equal(orc.stat[0][:, live], want['stat'][:, live]), p
            assert np.array_equal(_meta_rows(orc.meta[0], live), _meta_rows(want['meta'], live)), p
            vlive = util.live_vertex_mask(prog, want['cnt'], want['meta'])
            assert np.array_equal(orc.vtx[0][vlive], want['vtx'][vlive]), p
    finally:
        Oracle.force_factors(None)
        Oracle.set_sample_resets(False)


def _weighted_config():
    """CreateSprites with a random number of sprites per call and a Discrete factor with explicit
    probabilities (distributions.py:121-157, sprite_generators.py:75-77)."""
    import collections
    import moog_b200  # noqa: F401
    from moog import action_spaces, game_rules, physics as physics_lib, sprite, tasks
    from moog.state_initialization import distributions as distribs
    from moog.state_initialization import sprite_generators

    def state_initializer():
        return collections.OrderedDict([('agent', [sprite.Sprite(x=0.5, y=0.5, shape='square', scale=0.05)]), ('motes', [])])

    factors = distribs.Product(
        [distribs.Continuous('x', 0., 1.), distribs.Continuous('y', 0., 1.),
         distribs.Discrete('c0', [0.1, 0.5, 0.9], probs=[0.6, 0.3, 0.1]),
         distribs.Discrete('shape', ['triangle', 'square'], probs=[0.25, 0.75])],
        scale=0.02, c1=1., c2=1.)
    gen = sprite_generators.generate_sprites(factors, num_sprites=lambda: np.random.randint(1, 4))
    rules = (game_rules.CreateSprites('motes', gen),
             game_rules.VanishByFilter('motes', lambda s: s.x >= 0.))        # every mote lives for one pass
    return dict(state_initializer=state_initializer, physics=physics_lib.Physics(updates_per_env_step=1),
                task=tasks.CompositeTask(timeout_steps=1000), action_space=action_spaces.Grid(action_layers='agent'),
                observers={}, game_rules=rules)


def test_oracle_weighted_discrete_and_random_count():
    from moog_b200 import compiler
    from oracle.oracle import Oracle
    cfg = _weighted_config()
    states = [cfg['state_initializer']()]
    prog = compiler.compile_config(dict(cfg, game_rules=cfg['game_rules'][:1]), states, layer_capacity={'motes': 60})
    assert prog.rule_noise_dim == 1
    N = 64
    arrays = util.tile_state({k: compiler.pack_states(prog, states)[k] for k in util.STATE_KEYS}, N)
    orc = Oracle(prog, arrays)
    Oracle.set_seed(3)
    orc.post_reset()
    for t in range(14):
        Oracle.set_seed(20 + t)
        orc.step(np.full((N, 1), 4.0))
    lo = prog.layer_off[1]
    made = orc.cnt[:, 1]
    calls = 15 * N
    assert abs(made.sum() / calls - 2.0) < 5 * np.sqrt(2. / 3 / calls), made.sum() / calls        # uniform on {1, 2, 3}
    c0 = np.concatenate([orc.stat[e, 6, lo:lo + made[e]] for e in range(N)])
    shape = np.concatenate([orc.meta[e, 0, lo:lo + made[e]] for e in range(N)])
    n = len(c0)
    for value, p in ((0.1, 0.6), (0.5, 0.3), (0.9, 0.1)):
        assert abs((c0 == value).mean() - p) < 5 * np.sqrt(p * (1 - p) / n), (value, (c0 == value).mean())
    assert abs((shape == 0).mean() - 0.25) < 5 * np.sqrt(0.1875 / n)
    assert (orc.envi[:, 2] == 0).all()


@pytest.mark.gpu
def test_cuda_weighted_discrete_and_random_count_match_oracle():
    from moog_b200 import compiler
    from moog_b200.batched_env import Engine
    from oracle.oracle import Oracle
    cfg = _weighted_config()
    states = [cfg['state_initializer']()]
    prog = compiler.compile_config(cfg, states, layer_capacity={'motes': 8})
    N = 512
    arrays = util.tile_state({k: compiler.pack_states(prog, states)[k] for k in util.STATE_KEYS}, N)
    orc, eng = Oracle(prog, arrays), Engine(prog, N, 'cuda:0', seed=13)
    eng.state.upload(arrays)
    Oracle.set_seed(13)
    orc.post_reset()
    eng.post_reset()
    for t in range(12):
        Oracle.set_seed(eng.call_seed())
        orc.step(np.full((N, 1), 4.0))
        eng.env_step(np.full((N, 1), 4.0), auto_reset=False)
        dev = eng.state.download()
        assert np.array_equal(dev['cnt'], orc.cnt) and np.array_equal(dev['envi'][:, :6], orc.envi[:, :6]), t
        # the motes were written and popped in the same pass: the slots still hold what was drawn
        lo = prog.layer_off[1]
        assert np.array_equal(dev['stat'][:, :, lo:lo + 3], orc.stat[:, :, lo:lo + 3]), t
        assert np.array_equal(dev['dyn'][:, :, lo:lo + 3], orc.dyn[:, :, lo:lo + 3]), t
