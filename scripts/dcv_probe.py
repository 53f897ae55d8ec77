"""MOOG_PROFILE_DCV build: where a directed_collision_vectors call spends its cycles."""
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
true = c[:, 1].sum()
print('true overlaps %d (= %d DCV calls); contained vertices per DCV call %.2f' % (true, 2 * true, c[:, 2].sum() / (2 * true)))
print('per DCV call: contains+setup %.0f cycles, stage A %.0f, stage B+C %.0f ; env cycles mean %.3g' % (
    c[:, 5].sum() / (2 * true), c[:, 6].sum() / (2 * true), c[:, 7].sum() / (2 * true), c[:, 4].mean()))
