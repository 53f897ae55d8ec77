"""Auto-reset parity ("what's missing" 3): `Environment.step` of an env whose
previous transition terminated runs `reset()` instead and ignores the action
(/root/reference/moog/environment.py:100-101, :82-96).  bench.py's timed region takes that branch
for ~1 % of the envs every step, so it is pinned here:

* tests/golden/episodes_*.npz -- the UNMODIFIED reference stepped through several episodes
  (oracle/gen_golden_episodes.py); the oracle (CPU, here) and the CUDA path (-m gpu) follow it:
  state, FIRST / MID / LAST, None(NaN) reward and discount on FIRST, overlap pair sets, frames;
* the termination-timing case of the reference's own test_simulation.py, restated in tests/kat.py;
* -m gpu: a batch of envs in different episode phases, CUDA vs oracle, random pool rows.
"""
import numpy as np
import pytest

from tests import kat, util

EPISODE_SCENES = ['sim_timing', 'colliding_predators', 'falling_balls20']
EXACT = {'sim_timing', 'falling_balls20'}      # no sin / cos of a non-zero angle on the path


def _load(name):
    g = dict(np.load(util.GOLDEN + '/episodes_' + name + '.npz'))
    g['program'] = util.ProgramStub(g['blob'], g['layer_names'])
    g['pool'] = {k: g['pool_' + k] for k in util.STATE_KEYS}
    return g


def _first_state(g):
    st = {k: g['pool'][k][0:1].copy() for k in util.STATE_KEYS}
    st['envi'][:] = 0
    return st


def _same(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def _check_outputs(g, t, reward, step_type, discount, counters, what):
    assert int(step_type) == int(g['step_type'][t]), (what, t, 'step_type')
    assert _same(np.float32(reward), np.float32(g['reward'][t])), (what, t, 'reward', reward, g['reward'][t])
    assert _same(np.float32(discount), np.float32(g['discount'][t])), (what, t, 'discount')
    assert int(counters[0]) == int(g['n_calls'][t]), (what, t, 'overlap call count')
    assert int(counters[1]) == int(g['n_true'][t]), (what, t, 'overlap true count')
    assert np.uint64(int(counters[3]) & 0xFFFFFFFFFFFFFFFF) == g['true_hash'][t], (what, t, 'overlap pair set')


def _meta_for_comparison(meta, dyn):
    """Sprite OBJECTS that a config shares between episodes (colliding_predators.py:58 builds its
    walls once, outside the state initializer) keep whatever the previous episode did to them.
    The only such trace on this path is the NumPy dtype of a wall's angle_vel: `angle_vel += dw`
    with an np.float64 zero (collisions.py:449-454, the asymmetric predators x walls entry) turns
    the python float 0.0 into np.float64(0.0).  A pool of packed initial states restarts from the
    python float.  The dtype kind of an angle_vel that IS zero never reaches any arithmetic
    (sprite.py:426-430 skips a falsy angle_vel, every product with it is an exact zero), so the
    kind bits of such sprites are left out of the comparison -- and only those."""
    meta = meta.copy()
    zero_w = dyn[5] == 0
    meta[1, zero_w] &= ~(3 << 2)         # MOOG_SF_ANGVEL_SHIFT
    return meta


def _check_state(g, t, arrays, tol, what):
    """arrays: batch-of-1 record after step t."""
    prog = g['program']
    assert np.array_equal(arrays['cnt'][0], g['cnt'][t]), (what, t, 'cnt')
    live = util.live_mask(prog, g['cnt'][t])
    err = util.rel_err(arrays['dyn'][0][:, live], g['dyn'][t][:, live])
    f = np.nonzero(g['full_steps'] == t)[0]
    if len(f):
        f = int(f[0])
        err = max(err, util.rel_err(arrays['stat'][0][:, live], g['stat'][f][:, live]))
        assert np.array_equal(util.canonical_meta(_meta_for_comparison(arrays['meta'][0], arrays['dyn'][0]), live),
                              util.canonical_meta(_meta_for_comparison(g['meta'][f], g['dyn'][t]), live)), (what, t, 'meta')
        vlive = util.live_vertex_mask(prog, g['cnt'][t], g['meta'][f])
        err = max(err, util.rel_err(arrays['vtx'][0][vlive], g['vtx'][f][vlive]))
    assert err <= tol, (what, t, err)


@pytest.mark.parametrize('scene', EPISODE_SCENES)
def test_oracle_follows_reference_through_auto_resets(scene):
    """Bit for bit, every step of every episode, the reset steps included."""
    from oracle.oracle import Oracle
    g = _load(scene)
    prog = g['program']
    orc = Oracle(prog, _first_state(g))
    orc.post_reset()
    pool = Oracle(prog, g['pool'])
    T = len(g['step_type'])
    assert (g['reset_index'] >= 0).sum() >= 2
    for t in range(T):
        # a row is only consumed when the env itself decides to reset: junk on the other steps
        ri = int(g['reset_index'][t]) if g['reset_index'][t] >= 0 else (t * 7) % pool.n
        reward, step_type, discount = orc.step_auto(g['actions'][t][None], pool, [ri])
        _check_outputs(g, t, reward[0], step_type[0], discount[0], orc.counters[0], scene)
        _check_state(g, t, orc.arrays(), 0.0, scene)
        assert int(orc.envi[0, 1]) == int(g['step_type'][t] == 2), 'reset_next_step'
    frames = g['frames']
    assert (g['step_type'][g['frame_steps'][1:]] == 0).any(), 'a frame of a freshly reset env is on record'


def test_oracle_reference_termination_timing_kat():
    """test_simulation.py:68-78 testStep, three episodes in a row through the auto-reset."""
    from moog_b200 import compiler
    from oracle.oracle import Oracle
    cfg = kat.simulation_timing_config()
    states = [cfg['state_initializer']() for _ in range(2)]
    prog = compiler.compile_config(cfg, states)
    arrays = compiler.pack_states(prog, states)
    orc = Oracle(prog, {k: arrays[k][0:1] for k in util.STATE_KEYS})
    orc.post_reset()
    pool = Oracle(prog, arrays)
    for episode in range(3):
        for i, a in enumerate(kat.SIM_ACTIONS):
            _, st, _ = orc.step_auto([[a]], pool, [1])
            assert (int(st[0]) == 2) == (i == len(kat.SIM_ACTIONS) - 1), (episode, i, int(st[0]))
        reward, st, discount = orc.step_auto([[0]], pool, [1])
        assert int(st[0]) == 0 and np.isnan(reward[0]) and np.isnan(discount[0]), 'the next step is a reset'
        assert int(orc.envi[0, 3]) == episode + 1 and int(orc.envi[0, 0]) == 0, 'episode / step counters'


# ---------------------------------------------------------------------------
# CUDA path
# ---------------------------------------------------------------------------
def _engine(prog, arrays, pool):
    from moog_b200.batched_env import Engine
    eng = Engine(prog, arrays['dyn'].shape[0], 'cuda:0')
    eng.state.upload(arrays)
    eng.set_pool(pool)
    return eng


@pytest.mark.gpu
@pytest.mark.parametrize('scene', EPISODE_SCENES)
def test_cuda_follows_reference_through_auto_resets(scene):
    """`moog_env_step` with a pool and explicit `reset_index` along the reference's multi-episode
    trajectory: outputs and overlap pair sets identical to the reference's every step; state
    bit-exact (EXACT scenes) or within 1e-5 with the device re-synchronised on the oracle --
    itself bit-exact on this trajectory, see the CPU test -- before every step."""
    import torch
    from oracle.oracle import Oracle
    g = _load(scene)
    prog = g['program']
    exact = scene in EXACT
    eng = _engine(prog, _first_state(g), g['pool'])
    eng.post_reset()
    orc = Oracle(prog, _first_state(g))
    orc.post_reset()
    pool = Oracle(prog, g['pool'])
    T = len(g['step_type'])
    frame_at = {int(t): i for i, t in enumerate(g['frame_steps'])}
    for t in range(T):
        ri = int(g['reset_index'][t]) if g['reset_index'][t] >= 0 else (t * 7) % pool.n
        if not exact:
            eng.state.upload({k: v.copy() for k, v in orc.arrays().items()})
        orc.step_auto(g['actions'][t][None], pool, [ri])
        eng.env_step(g['actions'][t][None], auto_reset=True, reset_index=[ri], want_counters=True,
                     frames=True if t in frame_at else None)
        c = eng.counters[0].cpu().numpy()
        _check_outputs(g, t, float(eng.reward[0]), int(eng.step_type[0]), float(eng.discount[0]), c, scene)
        dev = eng.state.download()
        _check_state(g, t, dev, 0.0 if exact else util.RTOL, scene)
        assert np.array_equal(dev['envi'][0, [0, 1, 3]], orc.envi[0, [0, 1, 3]]), (scene, t, 'step / reset / episode counters')
        if t in frame_at:
            torch.cuda.synchronize()
            got = eng.frames[0].cpu().numpy()
            bad = int((got != g['frames'][frame_at[t]]).sum())
            assert bad == 0, (scene, t, 'frame', bad)


@pytest.mark.gpu
@pytest.mark.parametrize('scene', ['falling_balls20', 'colliding_predators84'])
def test_cuda_auto_reset_matches_oracle_batched(scene):
    """256 envs in different episode phases (staggered step counters, so that resets happen at
    different steps), several episodes each, random pool rows: CUDA vs oracle every step --
    state, step_type, reward (NaN on FIRST), discount, overlap pair sets, episode counters."""
    import importlib
    import moog_b200  # noqa: F401
    from moog_b200 import compiler
    from oracle.oracle import Oracle
    cfg = importlib.import_module('moog_b200.configs.' + scene).get_config()
    np.random.seed(5)
    states = [cfg['state_initializer']() for _ in range(12)]
    prog = compiler.compile_config(cfg, states)
    pool_arrays = compiler.pack_states(prog, states)
    pool_arrays = {k: pool_arrays[k] for k in util.STATE_KEYS}
    N = 256
    rng = np.random.RandomState(9)
    arrays = {k: np.ascontiguousarray(pool_arrays[k][rng.randint(0, 12, size=N)]) for k in util.STATE_KEYS}
    exact = scene == 'falling_balls20'
    timeout = 100 if exact else 200
    orc = Oracle(prog, arrays)
    orc.post_reset()
    orc.envi[:, 0] = rng.randint(timeout - 40, timeout - 1, size=N)   # step counters: time-outs 1..40 steps away
    eng = _engine(prog, orc.arrays(), pool_arrays)
    pool = Oracle(prog, pool_arrays)
    ad = max(prog.action_dim, 1)
    firsts = 0
    for t in range(60):
        ri = rng.randint(-2, 14, size=N)        # out-of-range rows are clamped by both
        act = rng.uniform(-1, 1, size=(N, ad)) if not exact else np.zeros((N, ad))
        if not exact:
            eng.state.upload({k: v.copy() for k, v in orc.arrays().items()})
        r_ref, st_ref, d_ref = orc.step_auto(act, pool, ri)
        eng.env_step(act, auto_reset=True, reset_index=ri, want_counters=True)
        assert np.array_equal(eng.step_type.cpu().numpy(), st_ref), (scene, t)
        assert _same(eng.reward.cpu().numpy(), r_ref.astype(np.float32)), (scene, t, 'reward')
        assert _same(eng.discount.cpu().numpy(), d_ref.astype(np.float32)), (scene, t, 'discount')
        assert np.array_equal(eng.counters.cpu().numpy()[:, :4], orc.counters), (scene, t, 'overlap pair sets')
        dev = eng.state.download()
        assert np.array_equal(dev['cnt'], orc.cnt) and np.array_equal(dev['envi'][:, [0, 1, 3]], orc.envi[:, [0, 1, 3]])
        worst = 0.0
        for e in range(N):
            live = util.live_mask(prog, orc.cnt[e])
            worst = max(worst, util.rel_err(dev['dyn'][e][:, live], orc.dyn[e][:, live]),
                        util.rel_err(dev['stat'][e][:, live], orc.stat[e][:, live]))
            vlive = util.live_vertex_mask(prog, orc.cnt[e], orc.meta[e])
            worst = max(worst, util.rel_err(dev['vtx'][e][vlive], orc.vtx[e][vlive]))
        assert worst <= (0.0 if exact else util.RTOL), (scene, t, worst)
        firsts += int((st_ref == 0).sum())
    assert firsts >= N, 'every env went through at least one auto-reset on average'


@pytest.mark.gpu
def test_cuda_reference_termination_timing_and_simulation_kat():
    """test_simulation.py:68-128 on the device: testStep's timing through BatchedEnvironment.step,
    and testSimStepSimPop through BatchedSimulation.sim_step / sim_pop / step."""
    import torch
    import moog_b200  # noqa: F401
    from moog_b200.batched_env import BatchedEnvironment
    from moog_b200.env_wrappers import BatchedSimulation
    cfg = kat.simulation_timing_config()
    N = 3
    env = BatchedEnvironment(**cfg, num_envs=N, device='cuda:0', seed=0, pool_size=2)
    act = lambda a: torch.full((N, 1), float(a), dtype=torch.float64)
    last = lambda ts: bool((ts.step_type == 2).all())
    none_last = lambda ts: not bool((ts.step_type == 2).any())

    def run_to_reward(step, actions):
        for a in actions[:-1]:
            assert none_last(step(act(a)))
        assert last(step(act(actions[-1])))

    # testStep, twice in a row: the step after a termination is the reset (FIRST, reward None)
    env.reset()
    for _ in range(2):
        run_to_reward(env.step, kat.SIM_ACTIONS)
        ts = env.step(act(0))
        assert bool((ts.step_type == 0).all()) and bool(ts.reward.isnan().all()) and bool(ts.discount.isnan().all())
    # testSimStepSimPop
    sim = BatchedSimulation(env)
    sim.reset()
    for a in kat.SIM_INIT:
        assert none_last(sim.sim_step(act(a)))
    for i in kat.SIM_POP_0:
        sim.sim_pop(i)
    run_to_reward(sim.sim_step, kat.SIM_REWARD_0)
    assert sim.sim_step(act(0)) is None, 'no simulation across an episode boundary (simulation.py:73-75)'
    for i in kat.SIM_POP_1:
        sim.sim_pop(i)
    run_to_reward(sim.sim_step, kat.SIM_REWARD_1)
    run_to_reward(sim.step, kat.SIM_ACTIONS)      # the real steps start from the state before any simulated one
    assert sim.stack_depth == 0


@pytest.mark.gpu
def test_cuda_logger_writes_the_reference_wire_format(tmp_path):
    """`BatchedLoggingEnvironment` against tests/golden/logger_sim_timing.json, which is what the
    UNMODIFIED reference's LoggingEnvironment wrote for the same environment and actions
    (oracle/gen_golden_logger.py; moog/env_wrappers/logger.py:134-224): same files, same step
    records -- reward (None on FIRST), step_type, action, layers, the 15 factors of every sprite,
    vertices on the steps the reference logs them -- for every logged env of the batch."""
    import json
    import os
    import torch
    import moog_b200  # noqa: F401
    from moog_b200.batched_env import BatchedEnvironment
    from moog_b200.env_wrappers import BatchedLoggingEnvironment
    gold = json.load(open(os.path.join(util.GOLDEN, 'logger_sim_timing.json')))
    N = 5
    env = BatchedEnvironment(**kat.simulation_timing_config(), num_envs=N, device='cuda:0', seed=0, pool_size=2)
    log = BatchedLoggingEnvironment(env, log_dir=str(tmp_path), log_vertices='WHEN_NECESSARY', envs=(0, 3))
    act = lambda a: torch.full((N, 1), float(a), dtype=torch.float64)
    log.reset()
    for _ in range(3):
        for a in kat.SIM_ACTIONS:
            log.step(act(a))
        log.step(act(4))
    for n in (0, 3):
        d = os.path.join(log.log_dir, 'env_%d' % n)
        assert json.load(open(os.path.join(d, 'attributes.txt'))) == gold['attributes']
        assert open(os.path.join(d, 'description.txt')).read() == gold['description']
        files = sorted(f for f in os.listdir(d) if f.isdigit())
        assert files == ['%05d' % i for i in range(len(gold['episodes']))]
        for fn, ref_ep in zip(files, gold['episodes']):
            ep = json.load(open(os.path.join(d, fn)))
            assert len(ep) == len(ref_ep), fn
            for k, (step, ref) in enumerate(zip(ep, ref_ep)):
                assert [x[0] for x in step[:5]] == ['time', 'reward', 'step_type', 'action', 'meta_state']
                assert step[1] == ref[1] and step[2] == ref[2] and step[3] == ref[3] and step[4] == ref[4], (fn, k)
                assert [l[0] for l in step[5]] == [l[0] for l in ref[5]], 'layers'
                for layer, ref_layer in zip(step[5], ref[5]):
                    assert len(layer[1]) == len(ref_layer[1])
                    for sp, ref_sp in zip(layer[1], ref_layer[1]):
                        assert len(sp) == len(ref_sp), (fn, k, 'vertices are logged on the same steps')
                        assert sp[:14] == ref_sp[:14], (fn, k, sp[:14], ref_sp[:14])
                        assert sp[14] is None          # metadata
                        if len(sp) == 17:
                            assert np.allclose(np.array(sp[16]), np.array(ref_sp[16]), rtol=0, atol=1e-15)
