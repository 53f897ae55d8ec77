"""Where a substep's cycles go, per env class (diagnostic; needs a MOOG_PROFILE_PHASES build:
MOOG_PROFILE_PHASES=1 MOOG_B200_LIB=<path> python -m moog_b200.build, then run with the same MOOG_B200_LIB).
counters: [4] total cycles, [5] refresh_candidates, [6] forces + collisions + correctives, [7] integrate_all."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from moog_b200.batched_env import BatchedEnvironment

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
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
for rep in range(2):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); eng.env_step(act, want_counters=True); e1.record(); torch.cuda.synchronize()
    c = eng.counters.cpu().numpy().astype(np.float64)
    tot, ref, frc, integ, true = c[:, 4], c[:, 5], c[:, 6], c[:, 7], c[:, 1]
    rest = tot - ref - frc - integ
    print('step %.3f ms; cycles per env: total %.3g = candidates %.3g + forces/collisions %.3g + integrate %.3g + rest (load, rules, task, store) %.3g' % (
        e0.elapsed_time(e1), tot.mean(), ref.mean(), frc.mean(), integ.mean(), rest.mean()))
    for lo, hi in ((0, 1), (1, 20), (20, 100), (100, 300), (300, 100000)):
        m = (true >= lo) & (true < hi)
        if m.any():
            print('  overlapping pairs in [%d,%d): %5d envs | total %.3g | candidates %.3g forces/coll %.3g integrate %.3g rest %.3g | share of all cycles %.1f%%' % (
                lo, hi, m.sum(), tot[m].mean(), ref[m].mean(), frc[m].mean(), integ[m].mean(), rest[m].mean(), 100 * tot[m].sum() / tot.sum()))
