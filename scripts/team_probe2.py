"""Team mode on the bench's own heaviest envs: the 4096-env phase mix of bench.py is stepped once
to find the envs that cost the most, their states are copied into batches of one wave each and
stepped in every launch mode.  With a MOOG_PROFILE_TEAM build (MOOG_B200_LIB=...) the counters
show how much of an env's time the team entries take and how wide they really run. (diagnostic)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from moog_b200.batched_env import BatchedEnvironment, Engine

E = 4096
cfg = bench._scene_config('falling_balls20')
states = bench._host_states(cfg, 256, 1234)
env = BatchedEnvironment(**cfg, num_envs=E, device='cuda:0', seed=1234, initial_states=states)
eng = env.engine
g = torch.Generator(device='cpu').manual_seed(1234)
act = torch.randint(0, 5, (E, env.action_dim), generator=g).to(torch.float64).to('cuda:0')
env.reset()
phase = torch.randint(0, 100, (E,), generator=g).to('cuda:0')
for t in range(130):
    if t < 100:
        eng.state.envi[:, 1] = torch.where(phase == t, torch.ones_like(phase, dtype=torch.int32), eng.state.envi[:, 1])
    eng.env_step(act)
snap = eng.state.download()
eng.env_step(act, auto_reset=False, want_counters=True)
torch.cuda.synchronize()
cyc = eng.counters.cpu().numpy()[:, 4].astype(np.float64)
order = np.argsort(-cyc)
profile = bool(os.environ.get('MOOG_PROFILE_TEAM'))
print('bench batch: cycles/env mean %.3g p50 %.3g p90 %.3g p99 %.3g max %.3g' % (
    cyc.mean(), *np.percentile(cyc, [50, 90, 99, 100])))
MODES = [('single', dict(MOOG_HELPER='0'), 10), ('helper', dict(MOOG_HELPER='1'), 5),
         ('team2', dict(MOOG_TEAM='2'), 6), ('team3', dict(MOOG_TEAM='3'), 4), ('team4', dict(MOOG_TEAM='4'), 3)]
for label, sel in (('heaviest 148', order[:148]), ('ranks 148..592', order[148:592]), ('median 444', order[E // 2 - 222:E // 2 + 222])):
    for name, envv, R in MODES:
        for k in ('MOOG_HELPER', 'MOOG_TEAM'):
            os.environ.pop(k, None)
        os.environ.update(envv)
        os.environ['MOOG_CTAS_PER_SM'] = str(R)
        n = 148 * R
        idx = np.resize(sel, n)
        arrays = {k: np.ascontiguousarray(snap[k][idx]) for k in snap}
        e2 = Engine(env.program, n, 'cuda:0')
        best = None
        for rep in range(3):
            e2.state.upload(arrays)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            e2.env_step(None, auto_reset=False, want_counters=True)
            e1.record()
            torch.cuda.synchronize()
            c = e2.counters.cpu().numpy().astype(np.float64)
            row = (e0.elapsed_time(e1), c)
            best = row if best is None or row[0] < best[0] else best
        ms, c = best
        line = '%-15s %-7s %2d envs/SM: launch %.3f ms | cycles/env mean %.3g max %.3g' % (label, name, R, ms, c[:, 4].mean(), c[:, 4].max())
        if profile and name.startswith('team'):
            line += ' | team entries/env %.1f: wall %.3g busy %.3g (x%.2f) wait %.3g | reference-order entries %.3g cycles' % (
                c[:, 0].mean(), c[:, 5].mean(), c[:, 6].mean(), c[:, 6].sum() / max(c[:, 5].sum(), 1), c[:, 7].mean(), c[:, 1].mean())
        print(line)
