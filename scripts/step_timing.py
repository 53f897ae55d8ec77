"""Per-step device time and contact counters of a scene (diagnostic)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from moog_b200.batched_env import BatchedEnvironment

scene = sys.argv[1] if len(sys.argv) > 1 else 'falling_balls20'
E = int(sys.argv[2]) if len(sys.argv) > 2 else 512
T = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cfg = bench._scene_config(scene)
POOL = int(os.environ.get('POOL', '64'))
states = bench._host_states(cfg, POOL, 1234)
env = BatchedEnvironment(**cfg, num_envs=E, device='cuda:0', seed=1, initial_states=states)
env.reset()
eng = env.engine
act = torch.zeros((E, env.action_dim), dtype=torch.float64, device='cuda:0')
if len(sys.argv) > 4 and sys.argv[4] == 'rand':
    act = torch.randint(0, 5, (E, env.action_dim)).to(torch.float64).to('cuda:0')
for t in range(T):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); eng.env_step(act, want_counters=True); e1.record(); torch.cuda.synchronize()
    c = eng.counters.double().mean(0).tolist()
    cm = eng.counters.double().max(0).values.tolist()
    top = int(eng.counters[:, 4].argmax())
    tc = eng.counters[top].tolist()
    print('step %3d  %8.3f ms  calls %.0f true %.1f coll %.1f  last %d | cycles mean %.3g max %.3g | narrow n %.1f cyc %.3g | resolve cyc %.3g' % (
        t, e0.elapsed_time(e1), c[0], c[1], c[2], int((eng.step_type == 2).sum()), c[4], cm[4], c[5], c[6], c[7]), 'top env', top, tc, flush=True)
