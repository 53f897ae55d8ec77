"""Shared helpers of the test-suite: golden fixtures and comparison rules."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

SCENES = ['pong', 'falling_balls', 'falling_balls20', 'colliding_predators',
          'predators_arena', 'synthetic32', 'falling_balls20_nan', 'cleanup',
          'chase_avoid_torus', 'pacman', 'timed_center', 'parallelogram_catch', 'forces_zoo',
          'reshape_zoo', 'portal_zoo', 'red_green', 'predict_zoo', 'bounce_box', 'functional_maze', 'multi_tracking', 'match_to_sample']
# scenes whose step() uses no sin/cos of a non-zero angle: every operation on
# the path is IEEE-exact (+ - * / sqrt fma), so the CUDA path must be bit-exact
EXACT_SCENES = ['pong', 'falling_balls', 'falling_balls20', 'falling_balls20_nan']

STATE_KEYS = ('dyn', 'stat', 'meta', 'vtx', 'cnt', 'envi', 'envf')

# north_star: single-step positions / velocities / angles within 1e-5 relative.
RTOL = 1e-5
ATOL_FLOOR = 1e-9   # absolute floor for components that are exactly 0 in the reference


class ProgramStub(object):
    """Just enough of moog_b200.compiler.Program, rebuilt from a stored blob."""

    def __init__(self, blob, layer_names=None):
        from moog_b200 import compiler as C
        self.blob = bytes(bytearray(blob))
        hdr = np.frombuffer(self.blob[:C.HDR_WORDS * 4], dtype='<i4')
        self.header = hdr
        self.n_layers = int(hdr[C.H_N_LAYERS])
        self.n_slots = int(hdr[C.H_N_SLOTS])
        self.K = int(hdr[C.H_K])
        self.n_envf = int(hdr[C.H_N_ENVF])
        self.action_dim = int(hdr[C.H_ACTION_DIM])
        self.noise_dim = int(hdr[C.H_NOISE_DIM])
        self.rule_noise_dim = int(hdr[C.H_RULE_NOISE_DIM])
        self.n_vtx = int(hdr[C.H_N_VTX])
        self.layer_off = [int(v) for v in hdr[C.H_LAYER_OFF:C.H_LAYER_OFF + self.n_layers + 1]]
        self.layer_cap = [b - a for a, b in zip(self.layer_off[:-1], self.layer_off[1:])]
        self.layer_names = [str(n) for n in layer_names] if layer_names is not None else [
            'layer%d' % i for i in range(self.n_layers)]
        self.render = None
        if hdr[C.H_R_ENABLED]:
            self.render = dict(height=int(hdr[C.H_R_HEIGHT]), width=int(hdr[C.H_R_WIDTH]),
                               aa=int(hdr[C.H_R_AA]))

    def layer_index(self, name):
        return self.layer_names.index(name)


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    g['program'] = ProgramStub(g['blob'], g['layer_names'])
    return g


def rule_noise_at(g, t):
    """[1, rule_noise_dim] uniforms behind ModifySprites(sample_one) at step t, or None."""
    prog = g['program']
    if not prog.rule_noise_dim or 'rule_noise' not in g:
        return None
    return np.asarray(g['rule_noise'][t], dtype=np.float64)[None, :prog.rule_noise_dim]


def reset_rule_noise(g):
    """[1, rule_noise_dim] uniforms behind the draws Environment.reset() made in the rules' reset
    (a Phase duration drawn with np.random.randint), or None."""
    if 'reset_rule_noise' not in g:
        return None
    return np.asarray(g['reset_rule_noise'], dtype=np.float64)[None, :g['program'].rule_noise_dim]


def state_at(g, t, prefix=None):
    """State record (batch of 1) after step t of the golden trajectory
    (t = -1: after reset).  envi / envf are not recorded per step."""
    if prefix is not None:
        return {k: g[prefix + '_' + k][None].copy() for k in STATE_KEYS}
    if t < 0:
        return {k: g['reset_' + k][None].copy() for k in STATE_KEYS}
    out = {k: g[k][t][None].copy() for k in ('dyn', 'stat', 'meta', 'vtx', 'cnt')}
    out['envi'] = g['reset_envi'][None].copy()
    out['envf'] = g['reset_envf'][None].copy()
    return out


def tile_state(st, n):
    return {k: np.repeat(v, n, axis=0) for k, v in st.items()}


def live_mask(prog, cnt):
    """[S] bool: slots that hold a live sprite."""
    m = np.zeros(prog.n_slots, dtype=bool)
    for l in range(prog.n_layers):
        m[prog.layer_off[l]:prog.layer_off[l] + int(cnt[l])] = True
    return m


def prog_voff(prog):
    """voff[S+1] (first cached vertex of each slot) read back from the program blob."""
    from moog_b200 import compiler as C
    blob = np.frombuffer(prog.blob, dtype=np.uint8)
    hdr = np.frombuffer(prog.blob[:C.HDR_WORDS * 4], dtype='<i4')
    n_ops = int(hdr[C.H_N_OPS])
    start = C.HDR_WORDS * 4 + 80 * n_ops
    ipool = np.frombuffer(blob[start:start + 4 * int(hdr[C.H_N_IPOOL])].tobytes(), dtype='<i4')
    return ipool[int(hdr[C.H_VOFF]):int(hdr[C.H_VOFF]) + prog.n_slots + 1]


def live_vertex_mask(prog, cnt, meta):
    """[VT] bool: cached vertices that belong to a live sprite (a slot freed by
    VanishOnContact keeps stale data on the device; the fixtures hold zeros)."""
    live = live_mask(prog, cnt)
    voff = prog_voff(prog)
    vlive = np.zeros(max(prog.n_vtx, 1), dtype=bool)
    for s in np.nonzero(live)[0]:
        vlive[voff[s]:voff[s] + int(meta[2, s])] = True
    return vlive


def canonical_meta(meta, live):
    """meta[3, S] of the live slots with the velocity-alias ids (MOOG_SF_VALIAS_SHIFT) renumbered
    1, 2, ... in slot order: the ids only name groups of sprites sharing one velocity ndarray
    (tether_physics.py:86-91); which number a group carries is an implementation detail (the
    device hands out fresh ids every substep, a packed reference state numbers them per state)."""
    out = np.array(meta[:, live], dtype=np.int64)
    out[1] &= ~64          # MOOG_SF_TELEPORTING: Portal's own bookkeeping (portal.py:36-76), not a sprite attribute
    ids = (out[1] >> 8) & 0x7fffff
    remap, nxt = {0: 0}, 0
    for k, i in enumerate(ids):
        if int(i) not in remap:
            nxt += 1
            remap[int(i)] = nxt
        out[1, k] = (out[1, k] & 0xff) | (remap[int(i)] << 8)
    # a group of one is no sharing at all
    new_ids = (out[1] >> 8) & 0x7fffff
    for i in set(new_ids.tolist()) - {0}:
        if (new_ids == i).sum() == 1:
            out[1, new_ids == i] &= 0xff
    return out


def assert_live_equal(prog, got, want, what):
    """Bit-equality of two state records (dicts of [fields, S] arrays / vtx /
    cnt) over the live sprites."""
    assert np.array_equal(got['cnt'], want['cnt']), what + ' cnt'
    live = live_mask(prog, want['cnt'])
    for k in ('dyn', 'stat'):
        assert np.array_equal(got[k][:, live], want[k][:, live]), what + ' ' + k
    assert np.array_equal(canonical_meta(got['meta'], live), canonical_meta(want['meta'], live)), what + ' meta'
    vlive = live_vertex_mask(prog, want['cnt'], want['meta'])
    assert np.array_equal(got['vtx'][vlive], want['vtx'][vlive]), what + ' vtx'


def rel_err(a, b):
    """max |a-b| / max(|b|, floor-scaled) as one number (0 when bit-equal)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    same = (a == b) | (np.isnan(a) & np.isnan(b))      # equal values (infinities included) or NaN on both sides
    finite = np.isfinite(a) & np.isfinite(b)
    if np.any(~same & ~finite):
        return float('inf')                            # NaN / inf on one side only, or infinities of opposite sign
    with np.errstate(invalid='ignore', over='ignore'):
        d = np.abs(np.where(finite, a, 0.0) - np.where(finite, b, 0.0))
        scale = np.maximum(np.abs(np.where(finite, b, 0.0)), ATOL_FLOOR / RTOL)
        e = np.where(same, 0.0, d / scale)
    return float(e.max())


AA_SCENES = ['pong', 'falling_balls20', 'colliding_predators', 'cleanup']


def with_anti_aliasing(g, aa):
    """The scene's program with PILRenderer(anti_aliasing=aa): the header word
    MOOG_H_R_AA of the stored blob is patched (compiler.py writes the same word
    from `renderer._anti_aliasing`)."""
    from moog_b200 import compiler as C
    blob = np.frombuffer(bytes(bytearray(g['blob'])), dtype=np.uint8).copy()
    hdr = blob[:C.HDR_WORDS * 4].view('<i4')
    hdr[C.H_R_AA] = aa
    return ProgramStub(blob, g['layer_names'])


def load_golden_aa(name):
    return dict(np.load(os.path.join(GOLDEN, name + '_aa.npz')))


BIG_SCENES = ['pong', 'colliding_predators']
BIG_SIZES = [(256, 256), (512, 512), (136, 200), (1024, 1024)]       # (width, height) as PILRenderer(image_size=...) takes them


def with_image_size(g, width, height):
    """The scene's program with PILRenderer(image_size=(width, height)): header words
    MOOG_H_R_WIDTH / MOOG_H_R_HEIGHT of the stored blob patched (compiler.py writes them from
    `renderer._image_size`)."""
    from moog_b200 import compiler as C
    blob = np.frombuffer(bytes(bytearray(g['blob'])), dtype=np.uint8).copy()
    hdr = blob[:C.HDR_WORDS * 4].view('<i4')
    hdr[C.H_R_WIDTH] = width
    hdr[C.H_R_HEIGHT] = height
    return ProgramStub(blob, g['layer_names'])


def load_golden_big(name):
    return dict(np.load(os.path.join(GOLDEN, name + '_big.npz')))


def device_reset_vs_oracle(env, exact):
    """`env.reset()` of a BatchedEnvironment(reset_mode='device') checked against the oracle, which draws
    the same generate_sprites groups from the same Philox stream (oracle/moog_oracle.c reset_generate):
    counts, dtype flags and outline sizes identical; attributes and outlines identical (`exact`: no
    sin / cos on the path) or within RTOL (a rotated sprite: device cos / sin vs libm).  Returns what
    env.reset() returned."""
    from oracle.oracle import Oracle
    eng, prog = env.engine, env.program
    seed = eng.call_seed()
    out = env.reset()
    dev = eng.state.download()
    pool_arrays = eng.pool.download()
    n = eng.n
    arrays = {k: np.repeat(pool_arrays[k][0:1], n, axis=0) for k in STATE_KEYS}
    arrays['envi'][:] = 0
    arrays['envi'][:, 1] = 1
    orc, pool = Oracle(prog, arrays), Oracle(prog, pool_arrays)
    Oracle.set_seed(seed)
    Oracle.set_sample_resets(True)
    try:
        orc.step_auto(None, pool, np.zeros(n, dtype=np.int32))
    finally:
        Oracle.set_sample_resets(False)
    assert np.array_equal(dev['cnt'], orc.cnt), 'device reset: counts'
    assert np.array_equal(dev['envi'][:, :6], orc.envi[:, :6]), 'device reset: counters / error words'
    worst = 0.0
    for e in range(n):
        live = live_mask(prog, orc.cnt[e])
        assert np.array_equal(canonical_meta(dev['meta'][e], live), canonical_meta(orc.meta[e], live)), (e, 'meta')
        vlive = live_vertex_mask(prog, orc.cnt[e], orc.meta[e])
        worst = max(worst, rel_err(dev['dyn'][e][:, live], orc.dyn[e][:, live]),
                    rel_err(dev['stat'][e][:, live], orc.stat[e][:, live]),
                    rel_err(dev['vtx'][e][vlive], orc.vtx[e][vlive]))
    assert worst <= (0.0 if exact else RTOL), ('device reset vs oracle', worst)
    return out
