"""Distribution of the per-env cycle counts of one env-step in the bench's phase mix (diagnostic)."""
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
for rep in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); eng.env_step(act, want_counters=True); e1.record(); torch.cuda.synchronize()
    c = eng.counters.cpu().numpy().astype(np.float64)
    cyc = c[:, 4]
    q = np.percentile(cyc, [50, 90, 99, 99.9, 100])
    print('step %.3f ms | cycles mean %.3g p50 %.3g p90 %.3g p99 %.3g p99.9 %.3g max %.3g | sum/1776 slots = %.3g cycles' % (
        e0.elapsed_time(e1), cyc.mean(), *q, cyc.sum() / 1776))
    pos = eng.state.dyn[:, 0:2, :20].cpu().numpy()
    far = (np.abs(pos - 0.5) > 1.0).any(axis=1)      # [E, 20] balls outside the arena
    nanb = np.isnan(pos).any(axis=1)
    print('envs with an escaped ball: %d, with a NaN ball: %d' % (far.any(axis=1).sum(), nanb.any(axis=1).sum()))
    top = np.argsort(-cyc)[:12]
    for i in top:
        print('  env %5d cycles %.3g  true %d coll %d narrow %d cyc_narrow %.3g cyc_resolve %.3g  escaped %d nan %d step_count %d' % (
            i, cyc[i], c[i, 1], c[i, 2], c[i, 5], c[i, 6], c[i, 7], far[i].sum(), nanb[i].sum(), int(eng.state.envi[i, 0])))
    # cost vs contacts
    coll = c[:, 2]
    for lo, hi in ((0, 1), (1, 10), (10, 30), (30, 60), (60, 100), (100, 200), (200, 1000)):
        m = (coll >= lo) & (coll < hi)
        if m.any():
            print('  coll in [%d,%d): %d envs, mean cycles %.3g, cycles/contact %.3g' % (lo, hi, m.sum(), cyc[m].mean(), (c[m, 7].sum() / max(coll[m].sum(), 1))))
