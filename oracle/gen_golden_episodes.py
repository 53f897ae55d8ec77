"""Golden vectors for the AUTO-RESET path: the UNMODIFIED reference stepped through several
episodes (environment.py:98-126 incl. the reset branch :100-101 and reset() :82-96).

TEST INFRASTRUCTURE ONLY; runs in the build container only (needs /root/reference, imported
through oracle/shims).  Usage:

    python oracle/gen_golden_episodes.py [scene ...]    # writes tests/golden/episodes_*.npz

The reference's `state_initializer` is random; the batched environment keeps such results in a
pool and hands a resetting env one row of it.  Here P initial states are drawn up front (packed
before anything mutates them -> `pool_*`), and the reference's initializer is replaced by one that
returns them in order, so that step t's reset uses pool row `reset_index[t]`.

Recorded per scene:
    blob, layer_names     program compiled from the reference's own config objects
    pool_*[P, ...]        packed initial states
    actions[T, A]         action fed at step t (ignored by the reference on a reset step)
    reset_index[T]        pool row the reset of step t took, -1 when step t was a normal step
    step_type[T]          0 FIRST (this step was a reset), 1 MID, 2 LAST
    reward[T], discount[T]   NaN where the reference returns None (FIRST)
    n_calls[T], n_true[T], true_hash[T]   Sprite.overlaps_sprite calls of step t
    dyn[T], cnt[T]        after every step;   stat / meta / vtx at `full_steps[F]`
    frames[G], frame_steps[G]   PILRenderer output (every FIRST step, the step before it, and some others)
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
from oracle.gen_golden import true_event_hash, _slot_map, _flat_action  # noqa: E402


def _sim_timing_config():
    """The environment of /root/reference/tests/moog/env_wrappers/test_simulation.py:35-63 without
    its meta_state rule (meta_state is host-side Python in MOOG)."""
    import collections
    from moog import action_spaces, observers, physics as physics_lib, sprite, tasks

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


_SIM_ACTIONS = [1, 4, 3, 1, 2, 0]   # test_simulation.py:72,101

SCENES = {
    # name: (config factory, seed, T, pool size, action policy)
    'sim_timing': (_sim_timing_config, 0, 24, 4, lambda env, t: _SIM_ACTIONS[t % 7] if t % 7 < 6 else 4),
    'colliding_predators': (
        lambda: importlib.import_module('moog_demos.example_configs.colliding_predators').get_config(None),
        21, 410, 4, None),       # episodes end by the 200-step timeout (ContactReward never resets here)
    'falling_balls20': (
        lambda: importlib.import_module('moog_b200.configs.falling_balls20').get_config(None),
        22, 215, 3, None),
}


def generate(name, out_dir):
    factory, seed, T, P, policy = SCENES[name]
    np.random.seed(seed)
    config = factory()
    env = environment.Environment(**config)
    initializer = env.state_initializer
    pool_states = [initializer() for _ in range(P)]
    prog = compiler.compile_config(config, pool_states)
    pool = compiler.pack_states(prog, pool_states)
    table = pool['shape_table']
    served = []

    def _next_state():
        assert len(served) < P, 'pool exhausted: raise P for scene ' + name
        served.append(len(served))
        return pool_states[served[-1]]

    env.state_initializer = _next_state
    renderer = config.get('observers', {}).get('image')
    ts = env.reset()
    assert served == [0]
    rec = {k: [] for k in ('actions', 'reset_index', 'step_type', 'reward', 'discount', 'n_calls', 'n_true',
                           'true_hash', 'dyn', 'cnt')}
    full = {k: [] for k in ('stat', 'meta', 'vtx')}
    full_steps, frames, frame_steps = [], [], []
    if renderer is not None:
        frames.append(np.asarray(ts.observation['image']))
        frame_steps.append(-1)
    prev_last = False
    for t in range(T):
        action = policy(env, t) if policy is not None else env.action_space.random_action()
        flat = _flat_action(prog, action)
        n_served = len(served)
        with refenv.OverlapLog() as log:
            ts = env.step(action)
        slots = _slot_map(prog, env.state)      # (a reset step: the new episode's sprites)
        h, n_true = 0, 0
        for a, b, r in log.calls:
            if r:
                n_true += 1
                h = true_event_hash(h, slots[id(a)], slots[id(b)])
        was_reset = len(served) > n_served
        assert was_reset == bool(ts.first()) == prev_last
        st = compiler.pack_states(prog, [env.state], table)
        rec['actions'].append(flat)
        rec['reset_index'].append(served[-1] if was_reset else -1)
        rec['step_type'].append(int(ts.step_type.value))
        rec['reward'].append(np.nan if ts.reward is None else float(ts.reward))
        rec['discount'].append(np.nan if ts.discount is None else float(ts.discount))
        rec['n_calls'].append(len(log.calls))
        rec['n_true'].append(n_true)
        rec['true_hash'].append(np.uint64(h))
        rec['dyn'].append(st['dyn'][0])
        rec['cnt'].append(st['cnt'][0])
        boundary = was_reset or ts.last()
        if boundary or t % 10 == 0 or t == T - 1:
            full_steps.append(t)
            for k in full:
                full[k].append(st[k][0])
        if renderer is not None and (boundary or t % 25 == 0):
            frames.append(np.asarray(ts.observation['image']))
            frame_steps.append(t)
        prev_last = bool(ts.last())
    n_resets = sum(1 for r in rec['reset_index'] if r >= 0)
    assert n_resets >= 2, '{}: only {} auto-resets in {} steps'.format(name, n_resets, T)
    out = dict(blob=np.frombuffer(prog.blob, dtype=np.uint8), layer_names=np.array(prog.layer_names),
               full_steps=np.array(full_steps, dtype=np.int32),
               frames=np.array(frames, dtype=np.uint8), frame_steps=np.array(frame_steps, dtype=np.int32))
    for k in ('dyn', 'stat', 'meta', 'vtx', 'cnt', 'envi', 'envf'):
        out['pool_' + k] = pool[k]
    for k, v in rec.items():
        out[k] = np.array(v)
    for k, v in full.items():
        out[k] = np.array(v)
    path = os.path.join(out_dir, 'episodes_' + name + '.npz')
    np.savez_compressed(path, **out)
    print('{:22s} T={:3d} auto-resets={} -> {} ({} KB)'.format(name, T, n_resets, path, os.path.getsize(path) // 1024))


def main():
    out_dir = os.path.join(_ROOT, 'tests', 'golden')
    for n in (sys.argv[1:] or list(SCENES)):
        generate(n, out_dir)


if __name__ == '__main__':
    main()
