"""Team mode (rows of a Collision entry side by side on W warps of one CTA): cycles of light /
medium / heavy falling_balls20 envs per launch mode, one wave of envs, every SM equally loaded.
(diagnostic)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from moog_b200.batched_env import Engine
from tests import util

g = util.load_golden('falling_balls20')
prog = g['program']
T = len(g['reward'])
parts = [util.state_at(g, t) for t in range(5, T - 1)]
probes = {'light(t=8)': 8 - 5, 'medium(t=24)': 24 - 5, 'heavy(t=38)': 38 - 5, 'heavy(t=43)': 43 - 5}
MODES = [('single', dict(MOOG_HELPER='0'), 10), ('helper', dict(MOOG_HELPER='1'), 5),
         ('team2', dict(MOOG_TEAM='2'), 6), ('team4', dict(MOOG_TEAM='4'), 3), ('team3', dict(MOOG_TEAM='3'), 4)]
for name, env, R in MODES:
    for k in ('MOOG_HELPER', 'MOOG_TEAM'):
        os.environ.pop(k, None)
    os.environ.update(env)
    os.environ['MOOG_CTAS_PER_SM'] = str(R)
    n = 148 * R
    idx = np.arange(n) % len(parts)
    arrays = {k: np.concatenate([parts[i][k] for i in idx], axis=0) for k in util.STATE_KEYS}
    eng = Engine(prog, n, 'cuda:0')
    best = None
    for rep in range(4):
        eng.state.upload(arrays)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.env_step(None, auto_reset=False, want_counters=True)
        e1.record()
        torch.cuda.synchronize()
        c = eng.counters.cpu().numpy()
        row = [e0.elapsed_time(e1), c[:, 4].mean()] + [c[idx == p, 4].mean() for p in probes.values()]
        best = row if best is None or row[0] < best[0] else best
    print('%-7s %2d envs/SM (%s): launch %.3f ms | cycles/env mean %.3g | ' % (
        name, R, eng.dev_program.step_launch_info(n), best[0], best[1]) + '  '.join(
        '%s %.3g' % (k, v) for k, v in zip(probes, best[2:])))
