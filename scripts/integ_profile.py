"""Cycles of the three phases of the Euler pass (MOOG_PROFILE_INTEG build): counters [5] phase 1
(lane = slot), [6] phase 2 (vertex cache), [7] phase 3 (boxes of rotated outlines). (diagnostic)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from moog_b200.batched_env import BatchedEnvironment
E = 4096
cfg = bench._scene_config('falling_balls20')
states = bench._host_states(cfg, 256, 1234)
env = BatchedEnvironment(**cfg, num_envs=E, device='cuda:0', seed=1234, initial_states=states)
eng = env.engine
g = torch.Generator(device='cpu').manual_seed(1234)
act = torch.randint(0, 5, (E, env.action_dim), generator=g).to(torch.float64).to('cuda:0')
env.reset()
for t in range(60):
    eng.env_step(act)
eng.env_step(act, want_counters=True); torch.cuda.synchronize()
c = eng.counters.cpu().numpy().astype(np.float64)
print('cycles per env-step: total %.3g | integrate phase 1 %.3g, phase 2 %.3g, phase 3 %.3g (per substep: %.0f / %.0f / %.0f)' % (
    c[:, 4].mean(), c[:, 5].mean(), c[:, 6].mean(), c[:, 7].mean(), c[:, 5].mean() / 20, c[:, 6].mean() / 20, c[:, 7].mean() / 20))
