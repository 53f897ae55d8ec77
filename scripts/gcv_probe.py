"""MOOG_PROFILE_GCV build: cycles inside _get_collision_vectors vs the rest of a handled overlap."""
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
act = torch.zeros((E, env.action_dim), dtype=torch.float64, device='cuda:0')
env.reset()
for t in range(45):
    eng.env_step(act)
eng.env_step(act, want_counters=True); torch.cuda.synchronize()
c = eng.counters.cpu().numpy().astype(np.float64)
n = c[:, 5].sum()
print('true overlaps %d resolved %d: owner DCV (incl. release barrier) %.0f cycles, then waiting for the helper %.0f cycles; env cycles mean %.3g max %.3g' % (
    n, c[:, 2].sum(), c[:, 6].sum() / n, c[:, 7].sum() / n, c[:, 4].mean(), c[:, 4].max()))
