"""Why does a warp run slower next to others when the schedulers are 88 % idle?  (diagnostic)

One wave of one-warp envs (148 x R envs, R resident per SM).  Batch A: every env holds the SAME
state (a contact-heavy falling_balls20 pile) -- all warps of an SM fetch the same instructions at
about the same time.  Batch B: envs hold different states of the same trajectory, the probe state
among them -- the warps of an SM are spread over the kernel's code.  If the slow-down with R comes
from the instruction caches, the probe state's cycles grow with R in batch B and much less in A.
"""
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
probe_t = int(sys.argv[1]) if len(sys.argv) > 1 else 38
probe = util.state_at(g, probe_t)
os.environ['MOOG_HELPER'] = '0'
for R in (1, 2, 4, 8, 12):
    os.environ['MOOG_CTAS_PER_SM'] = str(R)
    n = 148 * R
    out = {}
    for name in ('same', 'mixed'):
        if name == 'same':
            arrays = util.tile_state(probe, n)
            is_probe = np.ones(n, dtype=bool)
        else:
            parts = [util.state_at(g, t) for t in range(5, T - 1)]
            idx = np.arange(n) % len(parts)
            arrays = {k: np.concatenate([parts[i][k] for i in idx], axis=0) for k in util.STATE_KEYS}
            is_probe = idx == (probe_t - 5)
        eng = Engine(prog, n, 'cuda:0')
        cyc = []
        for rep in range(4):
            eng.state.upload(arrays)
            eng.env_step(None, auto_reset=False, want_counters=True)
            torch.cuda.synchronize()
            c = eng.counters.cpu().numpy()
            cyc.append(c[is_probe, 4].mean())
        out[name] = min(cyc[1:])
        info = eng.dev_program.step_launch_info(n)
    print('resident/SM %2d (launch info %s): probe env cycles  same-state batch %.3g   mixed batch %.3g' % (
        R, info, out['same'], out['mixed']))
