"""Pins the CPU oracle (oracle/moog_oracle.c, pil_oracle.c) against golden
vectors recorded from the UNMODIFIED reference (oracle/gen_golden.py).

Runs on CPU.  The oracle follows the golden trajectory step by step from the
recorded initial state, fed with the recorded actions (and the unit uniforms
behind RandomForce); after every step the full state record -- positions,
velocities, angles, the cached world vertices, sprite counts -- the reward,
the termination flag and the overlap-call statistics must match.
"""
import numpy as np
import pytest

from oracle.oracle import Oracle
from tests import util


@pytest.mark.parametrize('scene', util.SCENES)
def test_oracle_follows_reference_trajectory(scene):
    g = util.load_golden(scene)
    prog = g['program']
    orc = Oracle(prog, util.state_at(g, None, prefix='init'))
    orc.post_reset(rule_noise=util.reset_rule_noise(g))
    util.assert_live_equal(prog, {k: getattr(orc, k)[0] for k in ('dyn', 'stat', 'vtx', 'cnt', 'meta')},
                           {k: g['reset_' + k] for k in ('dyn', 'stat', 'vtx', 'cnt', 'meta')}, 'reset')
    T = len(g['reward'])
    worst = 0.0
    for t in range(T):
        noise = g['noise'][t][None] if prog.noise_dim else None
        reward, step_type = orc.step(g['actions'][t][None], noise=noise, rule_noise=util.rule_noise_at(g, t))
        assert np.array_equal(orc.cnt[0], g['cnt'][t]), (scene, t)
        live = util.live_mask(prog, g['cnt'][t])
        for k in ('dyn', 'stat'):
            worst = max(worst, util.rel_err(getattr(orc, k)[0][:, live], g[k][t][:, live]))
        vlive = util.live_vertex_mask(prog, g['cnt'][t], g['meta'][t])
        worst = max(worst, util.rel_err(orc.vtx[0][vlive], g['vtx'][t][vlive]))
        assert np.array_equal(util.canonical_meta(orc.meta[0], live), util.canonical_meta(g['meta'][t], live)), (scene, t, 'meta')
        assert reward[0] == g['reward'][t], (scene, t)
        assert bool(step_type[0] == 2) == bool(g['last'][t]), (scene, t)
        n_calls, n_true, _, h = orc.counters[0]
        assert n_calls == g['n_calls'][t], (scene, t, 'overlap call count')
        assert n_true == g['n_true'][t], (scene, t, 'overlap true count')
        assert np.uint64(h) == g['true_hash'][t], (scene, t, 'overlap pair set')
        assert worst == 0.0, (scene, t, worst)   # the oracle is bit-exact vs the reference


@pytest.mark.parametrize('scene', [s for s in util.SCENES])
def test_oracle_render_matches_reference_frames(scene):
    g = util.load_golden(scene)
    prog = g['program']
    if prog.render is None or len(g['frames']) == 0:
        pytest.skip('scene has no renderer')
    for f, t in zip(g['frames'], g['frame_steps']):
        orc = Oracle(prog, util.state_at(g, int(t)))
        out = orc.render()[0]
        assert out.shape == f.shape
        assert np.array_equal(out, f), (scene, int(t), int((out != f).sum()))


@pytest.mark.parametrize('aa', [2, 3])
@pytest.mark.parametrize('scene', util.AA_SCENES)
def test_oracle_render_matches_reference_antialiased_frames(scene, aa):
    """PILRenderer(anti_aliasing=aa): supersampled canvas + Image.resize(LANCZOS)
    (pil_renderer.py:65-66,113); frames recorded from the reference renderer."""
    g = util.load_golden(scene)
    ga = util.load_golden_aa(scene)
    prog = util.with_anti_aliasing(g, aa)
    for f, t in zip(ga['frames_aa%d' % aa], ga['frame_steps']):
        out = Oracle(prog, util.state_at(g, int(t))).render()[0]
        assert np.array_equal(out, f), (scene, aa, int(t), int((out != f).sum()))


@pytest.mark.parametrize('size', util.BIG_SIZES)
@pytest.mark.parametrize('scene', util.BIG_SCENES)
def test_oracle_render_matches_reference_big_frames(scene, size):
    """Canvases beyond one CTA's shared memory (the shipped pacman draws 256 x 256,
    tests/runtime_benchmark.py:31-38 times 256 .. 1024), and a non-square one: frames recorded from
    the reference renderer at those sizes."""
    g = util.load_golden(scene)
    gb = util.load_golden_big(scene)
    prog = util.with_image_size(g, *size)
    for f, t in zip(gb['frames_%dx%d' % size], gb['frame_steps']):
        out = Oracle(prog, util.state_at(g, int(t))).render()[0]
        assert out.shape == f.shape
        assert np.array_equal(out, f), (scene, size, int(t), int((out != f).sum()))


@pytest.mark.parametrize('scene', util.SCENES)
def test_rows_of_a_collision_entry_commute(scene):
    """The rule the next kernel design rests on ("Next" (1)), checked on the reference's
    own trajectories: with the oracle's row-order mode on, every Collision entry is executed in
    waves of mutually independent rows, each wave in REVERSE row order (oracle/moog_oracle.c
    `collision_op_rows`).  State, rewards and the overlap statistics -- the order-sensitive hash of
    the True pairs is folded from per-row logs in row order -- must still equal the golden
    trajectory bit for bit, and no sprite may outrun the margin the independence test allows."""
    import ctypes
    from oracle import oracle as orc_mod
    L = orc_mod.lib()
    L.orc_set_row_mode.argtypes = [ctypes.c_int]
    L.orc_row_stats.argtypes = [ctypes.c_void_p]
    L.orc_set_row_mode(1)
    try:
        test_oracle_follows_reference_trajectory(scene)
        stats = (ctypes.c_int64 * 4)()
        L.orc_row_stats(stats)
    finally:
        L.orc_set_row_mode(0)
    entries, rows, waves, outran = list(stats)
    if entries:
        # the schedule is not the sequential one (fewer waves than rows), and entries in which a
        # sprite outran its margin -- where a kernel would fall back to the reference order --
        # are the exception
        assert waves <= rows, (scene, entries, rows, waves)
        assert outran <= 0.05 * entries, (scene, entries, outran)
        print('%s: %d entries with several rows, %.2f rows per wave, %d sprites outran their margin' % (
            scene, entries, rows / waves, outran))


@pytest.mark.parametrize('name,floor', [('falling_balls20', 100), ('synthetic32', 0), ('colliding_predators84', 0)])
def test_rows_commute_in_heavy_piles(name, floor):
    """The same rule where it matters most (and, with synthetic32 / colliding_predators84, for
    symmetric entries that also exchange spin, under random actions): whole 100-step episodes of falling_balls20 (first
    impacts of 20 balls, settled piles with up to a few hundred overlapping pairs per env-step --
    far more than the golden trajectory reaches) stepped by the oracle in the reference's order
    and in the row-order mode from identical initial states must agree bit for bit after every
    step, overlap statistics and the order-sensitive pair hash included."""
    import ctypes
    import moog_b200  # noqa: F401
    from moog_b200 import compiler
    import importlib
    from oracle import oracle as orc_mod
    cfg = importlib.import_module('moog_b200.configs.' + name).get_config()
    np.random.seed(5)
    states = [cfg['state_initializer']() for _ in range(6)]
    prog = compiler.compile_config(cfg, states)
    arrays = compiler.pack_states(prog, states)
    L = orc_mod.lib()
    L.orc_set_row_mode.argtypes = [ctypes.c_int]
    L.orc_row_stats.argtypes = [ctypes.c_void_p]
    a, b = Oracle(prog, arrays), Oracle(prog, arrays)
    a.post_reset()
    b.post_reset()
    rng = np.random.RandomState(3)
    most = 0
    for t in range(100):
        actions = rng.uniform(-1, 1, size=(len(states), max(prog.action_dim, 1)))
        if name == 'falling_balls20':
            actions = np.zeros_like(actions)
        L.orc_set_row_mode(0)
        a.step(actions)
        L.orc_set_row_mode(1)
        try:
            b.step(actions)
        finally:
            L.orc_set_row_mode(0)
        for k in ('dyn', 'stat', 'vtx', 'cnt', 'meta'):
            assert np.array_equal(getattr(a, k), getattr(b, k), equal_nan=(k in ('dyn', 'stat', 'vtx'))), (t, k)
        assert np.array_equal(a.counters, b.counters), (t, 'overlap statistics / pair hash')
        most = max(most, int(a.counters[:, 1].max()))
    assert most >= floor, 'the episodes never reached a contact-heavy pile (%d)' % most
