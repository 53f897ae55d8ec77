"""compute-sanitizer target: the round-2 code paths on small batches (CreateSprites + Bernoulli
conditions with auto-resets, device-side reset sampling, metadata columns, decision trees, a traced
user-defined rule), each checked against the oracle as it goes.

    compute-sanitizer --tool memcheck python scripts/sanitize_new_paths.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import moog_b200  # noqa: E402,F401
from moog_b200 import compiler  # noqa: E402
from moog_b200.batched_env import BatchedEnvironment, Engine  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from tests import util  # noqa: E402


def spawn(n=12, steps=70):
    mod = importlib.import_module('moog_b200.configs.spawn_zoo')
    cfg = mod.get_config()
    np.random.seed(3)
    states = [cfg['state_initializer']() for _ in range(4)]
    prog = compiler.compile_config(cfg, states, layer_capacity={'drops': 6, 'sparks': 5})     # overflow on purpose
    pool = {k: v for k, v in compiler.pack_states(prog, states).items() if k in util.STATE_KEYS}
    rng = np.random.RandomState(1)
    arrays = {k: np.ascontiguousarray(pool[k][rng.randint(0, 4, size=n)]) for k in util.STATE_KEYS}
    orc, eng = Oracle(prog, arrays), Engine(prog, n, 'cuda:0', seed=5)
    eng.state.upload(arrays)
    eng.set_pool(pool)
    opool = Oracle(prog, pool)
    Oracle.set_seed(5)
    orc.post_reset()
    eng.post_reset()
    for t in range(steps):
        ri = rng.randint(0, 4, size=n)
        act = rng.uniform(-1, 1, size=(n, 2))
        Oracle.set_seed(eng.call_seed())
        orc.step_auto(act, opool, ri)
        eng.env_step(act, auto_reset=True, reset_index=ri, frames=True)
        dev = eng.state.download()
        assert np.array_equal(dev['cnt'], orc.cnt) and np.array_equal(dev['envi'][:, :6], orc.envi[:, :6]), t
    assert np.array_equal(eng.frames.cpu().numpy(), orc.render())
    print('spawn_zoo ok', orc.cnt[:, 1:3].max(axis=0), 'overflow flags', int((orc.envi[:, 2] & 8 != 0).sum()))


def golden(name, steps):
    g = util.load_golden(name)
    prog = g['program']
    n = 3
    eng = Engine(prog, n, 'cuda:0')
    eng.state.upload(util.tile_state(util.state_at(g, 0, prefix='init'), n))
    eng.post_reset()          # task / action / rule state armed on the device (TimedRule intervals, rule attributes)
    for t in range(min(steps, len(g['reward']))):
        act = np.repeat(g['actions'][t][None], n, axis=0)
        nz = np.repeat(g['noise'][t][None], n, axis=0).reshape(n, -1) if prog.noise_dim else None
        rn = np.repeat(g['rule_noise'][t][None], n, axis=0)[:, :max(prog.rule_noise_dim, 1)] if prog.rule_noise_dim else None
        eng.env_step(act, noise=nz, rule_noise=rn, auto_reset=False, frames=True if prog.render else None)
        dev = eng.state.download()
        assert np.array_equal(dev['cnt'][0], g['cnt'][t]), (name, t)
    print(name, 'ok')


def device_resets():
    from moog_b200.configs import colliding_predators84
    cfg = colliding_predators84.get_config()
    np.random.seed(2)
    states = [cfg['state_initializer']() for _ in range(3)]
    env = BatchedEnvironment(**cfg, num_envs=16, device='cuda:0', seed=4, initial_states=states, reset_mode='device')
    util.device_reset_vs_oracle(env, exact=False)
    print('device resets ok')


if __name__ == '__main__':
    spawn()
    for name, steps in (('bounce_box', 32), ('red_green', 115), ('functional_maze', 70), ('portal_zoo', 20)):
        golden(name, steps)
    device_resets()
    print('ALL OK')
