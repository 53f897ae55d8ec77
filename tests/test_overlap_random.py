"""Randomised narrow-phase parity (SURVEY section 7 step 3 / hard part 2; VERDICT round 1 "weak" 3):
`moog_overlap_pairs` -- the CUDA `overlaps()` with its padded-box culls, edge-box prefilter and
pair-parallel `segments_intersect` (csrc/moog_step.cu path_intersects_filled_impl) -- against the
oracle's plain restatement of `Sprite.overlaps_sprite` / matplotlib `path_intersects_path`
(oracle/moog_oracle.c), on >= 10^5 seeded sprite pairs per shape class.

Every env holds 8 sprites in layer `a` and 8 in layer `b`; all 64 (a, b) pairs are tested.  The
pairs (a_j, b_j) are *designed*: b_j is translated into a special relation to a_j --

    1  centre distance around the sum of the circumscribed radii (the broad-phase threshold)
    2  a vertex of b exactly on a vertex of a            3  ... offset by +-1e-12 / 1e-9 / 1e-7
    4  a vertex of b on an edge of a                     5  ... offset along the edge normal
    6  an edge of b collinear with an edge of a (both outlines rotated to make them parallel)
    7  b's centre on a's centre (containment / identical outlines)

-- the other 56 pairs of the env are whatever the random placement gives.  The number of flips
(results that differ) is reported and must be zero: a culled pair is one the reference's
`overlaps_sprite` returns False for.
"""
import collections

import numpy as np
import pytest

from tests import util

CLASSES = {
    # name: (shapes of layer a, shapes of layer b, aspect-ratio range)
    'circle': (['circle'], ['circle'], (1.0, 1.0)),              # the 30-gon of shapes.py:20 (disc containment path)
    'ellipse': (['circle'], ['circle'], (0.6, 1.6)),
    'star': (['star_5', 'star_4', 'star_6'], ['star_5'], (0.8, 1.25)),
    'spoke': (['spoke_4', 'spoke_5', 'spoke_6'], ['spoke_4'], (0.8, 1.25)),
    'triangle': (['triangle'], ['triangle'], (0.5, 2.0)),
    'walls': (['square'], ['square', 'triangle', 'circle'], (0.2, 5.0)),   # long thin boxes vs everything
    'mixed': (['circle', 'star_5', 'spoke_4', 'triangle', 'square', 'pentagon', 'hexagon', 'star_6'],
              ['square', 'circle', 'spoke_6', 'star_4', 'triangle', 'hexagon', 'spoke_5', 'pentagon'], (0.7, 1.4)),
}
PER_LAYER = 8


def _program(cls):
    """A two-layer program (no physics entries needed) and one packed env."""
    import moog_b200  # noqa: F401
    from moog import action_spaces, physics as physics_lib, sprite, tasks
    from moog_b200 import compiler
    sa, sb, _ = CLASSES[cls]
    mk = lambda shape, k: sprite.Sprite(x=0.5, y=0.5, shape=shape, scale=0.1, c0=k)
    state = collections.OrderedDict([
        ('a', [mk(sa[k % len(sa)], k) for k in range(PER_LAYER)]),
        ('b', [mk(sb[k % len(sb)], k) for k in range(PER_LAYER)]),
        ('agent', [])])
    cfg = dict(state_initializer=lambda: state, physics=physics_lib.Physics(updates_per_env_step=1),
               task=tasks.CompositeTask(timeout_steps=10), action_space=action_spaces.Grid(action_layers='agent'),
               observers={}, game_rules=())
    prog = compiler.compile_config(cfg, [state])
    arrays = compiler.pack_states(prog, [state])
    return prog, {k: arrays[k] for k in util.STATE_KEYS}


def build_batch(cls, n_envs, seed):
    """-> (program, arrays of n_envs envs, relation[n_envs, 8])."""
    prog, one = _program(cls)
    rng = np.random.RandomState(seed)
    S = prog.n_slots
    voff = util.prog_voff(prog)
    nv = one['meta'][0, 2]
    arr = {k: np.repeat(v, n_envs, axis=0) for k, v in one.items()}
    lo_a, lo_b = prog.layer_off[prog.layer_index('a')], prog.layer_off[prog.layer_index('b')]
    lo_aspect, hi_aspect = CLASSES[cls][2]
    pos = np.zeros((n_envs, S, 2))
    # random similarity (+ aspect) transform of every sprite about its centre
    for s in list(range(lo_a, lo_a + PER_LAYER)) + list(range(lo_b, lo_b + PER_LAYER)):
        v = one['vtx'][0, voff[s]:voff[s] + nv[s]] - one['dyn'][0, 0:2, s]           # outline about the centre
        scale = rng.uniform(0.35, 1.6, size=n_envs)
        aspect = np.exp(rng.uniform(np.log(lo_aspect), np.log(hi_aspect), size=n_envs)) if hi_aspect > lo_aspect else np.ones(n_envs)
        ang = rng.uniform(0, 2 * np.pi, size=n_envs) if cls not in ('circle',) else np.zeros(n_envs)
        if cls == 'walls' and s < lo_a + PER_LAYER:
            ang = rng.randint(0, 4, size=n_envs) * (np.pi / 2) * (rng.uniform(size=n_envs) < 0.7)   # mostly axis-aligned
        x = v[None, :, 0] * scale[:, None]
        y = v[None, :, 1] * (scale * aspect)[:, None]
        c, sn = np.cos(ang)[:, None], np.sin(ang)[:, None]
        w = np.stack([c * x - sn * y, sn * x + c * y], axis=2)                         # [n, nv, 2]
        p = rng.uniform(0.15, 0.85, size=(n_envs, 2))
        pos[:, s] = p
        arr['vtx'][:, voff[s]:voff[s] + nv[s]] = w + p[:, None, :]
        arr['stat'][:, 5, s] = np.sqrt((w * w).sum(axis=2)).max(axis=1)               # circumscribed radius
        arr['stat'][:, 2, s] = aspect                                                  # is_symmetric_circle reads it
    relation = rng.randint(0, 8, size=(n_envs, PER_LAYER))
    for j in range(PER_LAYER):
        sa, sb = lo_a + j, lo_b + j
        A = arr['vtx'][:, voff[sa]:voff[sa] + nv[sa]]
        B = arr['vtx'][:, voff[sb]:voff[sb] + nv[sb]]
        rel = relation[:, j]
        e = np.arange(n_envs)
        ia, ib = rng.randint(0, nv[sa], size=n_envs), rng.randint(0, nv[sb], size=n_envs)
        pa, pa2 = A[e, ia], A[e, (ia + 1) % nv[sa]]
        qb = B[e, ib]
        t = np.zeros((n_envs, 2))
        # 1: around the broad-phase threshold
        d = rng.uniform(0, 2 * np.pi, size=n_envs)
        u = np.where(rng.uniform(size=n_envs) < 0.3, 1.0 + rng.choice([-1e-12, 0.0, 1e-12, 1e-9, -1e-9], size=n_envs),
                     rng.uniform(0.55, 1.08, size=n_envs))
        reach = (arr['stat'][:, 5, sa] + arr['stat'][:, 5, sb]) * u
        t1 = pos[:, sa] + reach[:, None] * np.stack([np.cos(d), np.sin(d)], 1) - pos[:, sb]
        # 2 / 3: vertex on vertex (+ offset)
        off = rng.choice([1e-12, -1e-12, 1e-9, -1e-9, 1e-7, -1e-7, 3e-7], size=(n_envs, 2))
        t2 = pa - qb
        t3 = t2 + off
        # 4 / 5: vertex on edge (+ offset along the normal)
        lam = rng.uniform(0, 1, size=n_envs)[:, None]
        on_edge = pa + lam * (pa2 - pa)
        nrm = np.stack([(pa2 - pa)[:, 1], -(pa2 - pa)[:, 0]], 1)
        nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-300)
        t4 = on_edge - qb
        t5 = t4 + nrm * rng.choice([1e-12, -1e-12, 1e-9, -1e-9, 1e-7, -1e-7], size=n_envs)[:, None]
        # 7: common centre
        t7 = pos[:, sa] - pos[:, sb] + rng.choice([0.0, 1e-9, 1e-3], size=n_envs)[:, None]
        for code, tt in ((1, t1), (2, t2), (3, t3), (4, t4), (5, t5), (7, t7)):
            t[rel == code] = tt[rel == code]
        # 6: collinear edges -- rotate b about its vertex so that its edge is parallel to a's, put the
        # vertex on a's edge, slide along it
        m6 = rel == 6
        if m6.any():
            qb2 = B[e, (ib + 1) % nv[sb]]
            ea, eb = pa2 - pa, qb2 - qb
            dth = np.arctan2(ea[:, 1], ea[:, 0]) - np.arctan2(eb[:, 1], eb[:, 0]) + np.pi * rng.randint(0, 2, size=n_envs)
            c, sn = np.cos(dth)[:, None, None], np.sin(dth)[:, None, None]
            rel_v = B - qb[:, None, :]
            rot = np.concatenate([c * rel_v[..., 0:1] - sn * rel_v[..., 1:2], sn * rel_v[..., 0:1] + c * rel_v[..., 1:2]], axis=2)
            slide = rng.uniform(-0.5, 1.2, size=n_envs)[:, None]
            newB = rot + (pa + slide * ea)[:, None, :]
            B[m6] = newB[m6]
            cpos = pos[:, sb] - qb
            cposr = np.stack([c[:, 0, 0] * cpos[:, 0] - sn[:, 0, 0] * cpos[:, 1], sn[:, 0, 0] * cpos[:, 0] + c[:, 0, 0] * cpos[:, 1]], 1)
            pos[m6, sb] = (cposr + pa + slide * ea)[m6]
        B += t[:, None, :]
        pos[:, sb] += t
    arr['dyn'][:, 0, :] = pos[:, :, 0]
    arr['dyn'][:, 1, :] = pos[:, :, 1]
    return prog, arr, relation


def test_oracle_overlap_agrees_with_host_geometry_on_designed_pairs():
    """CPU: the oracle's overlaps_sprite against the host library's independent C++ restatement
    (`moog_host_paths_overlap`, what the host `Sprite.overlaps_sprite` calls) on the designed
    pairs of every class -- two implementations of matplotlib's algorithm written apart."""
    import ctypes
    from moog_b200 import capi
    from oracle.oracle import Oracle
    L = capi.lib()
    for cls in CLASSES:
        prog, arr, _ = build_batch(cls, 96, seed=11)
        orc = Oracle(prog, arr)
        ref = orc.overlap_pairs('a', 'b')
        voff = util.prog_voff(prog)
        nv = arr['meta'][0, 2]
        lo_a, lo_b = prog.layer_off[prog.layer_index('a')], prog.layer_off[prog.layer_index('b')]
        flips = 0
        for e in range(96):
            for i in range(PER_LAYER):
                for j in range(PER_LAYER):
                    sa, sb = lo_a + i, lo_b + j
                    dist = np.sqrt(np.dot(arr['dyn'][e, 0:2, sa] - arr['dyn'][e, 0:2, sb], arr['dyn'][e, 0:2, sa] - arr['dyn'][e, 0:2, sb]))
                    if dist > arr['stat'][e, 5, sa] + arr['stat'][e, 5, sb]:
                        continue        # sprite.py:464-466 decides; the outline test is not reached
                    pa = arr['vtx'][e, voff[sa]:voff[sa] + nv[sa]]
                    pb = arr['vtx'][e, voff[sb]:voff[sb] + nv[sb]]
                    pa = np.ascontiguousarray(np.vstack([pa, pa[:1]]))
                    pb = np.ascontiguousarray(np.vstack([pb, pb[:1]]))
                    got = L.moog_host_paths_overlap(pa.ctypes.data_as(ctypes.c_void_p), len(pa),
                                                    pb.ctypes.data_as(ctypes.c_void_p), len(pb))
                    flips += bool(got) != bool(ref[e, i, j])
        assert flips == 0, (cls, flips)


@pytest.mark.gpu
@pytest.mark.parametrize('cls', sorted(CLASSES))
def test_cuda_overlap_pairs_match_oracle_on_random_pairs(cls):
    from moog_b200.batched_env import Engine
    from oracle.oracle import Oracle
    n_envs = 1664                                   # x 64 pairs = 106 496 pairs per class
    prog, arr, relation = build_batch(cls, n_envs, seed=2024)
    ref = Oracle(prog, arr).overlap_pairs('a', 'b')
    eng = Engine(prog, n_envs, 'cuda:0')
    eng.state.upload(arr)
    got = eng.overlap_pairs('a', 'b').cpu().numpy()
    flips = int((got != ref).sum())
    diag = np.arange(PER_LAYER)
    designed = ref[:, diag, diag]
    by_rel = {int(r): (int(designed[relation == r].sum()), int((relation == r).sum())) for r in range(8)}
    print('{}: {} pairs, {} overlapping, {} flips; designed pairs true/total by relation: {}'.format(
        cls, ref.size, int(ref.sum()), flips, by_rel))
    assert ref.size >= 100000
    assert 0.02 < ref.mean() < 0.9, 'degenerate test data'
    for r in (2, 3, 4, 5, 6):
        true, total = by_rel[r]
        assert 0 < true, (cls, r, 'the designed relation never overlaps')
    assert flips == 0, (cls, flips, np.argwhere(got != ref)[:5])
    # and the other argument order (overlaps_sprite is not symmetric in its arguments' roles)
    ref_ba = Oracle(prog, arr).overlap_pairs('b', 'a')
    got_ba = eng.overlap_pairs('b', 'a').cpu().numpy()
    assert int((got_ba != ref_ba).sum()) == 0, cls
