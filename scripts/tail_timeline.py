"""Timeline of one env-step of the bench batch with frames (diagnostic, MOOG_TRACE_TIMES=1): when each
env's step starts / finishes and when its frame is drawn, relative to the first step CTA."""
import sys, os
os.environ['MOOG_TRACE_TIMES'] = '1'
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
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda:0')
for rep in range(3):
    flush.fill_(rep)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); eng.env_step(act, want_counters=True, frames=True); e1.record(); torch.cuda.synchronize()
    c = eng.counters.cpu().numpy().astype(np.int64)
    s0, s1, r0, r1 = c[:, 6], c[:, 7], c[:, 4], c[:, 5]
    t0 = s0.min()
    ms = lambda x: (x - t0) / 1e6
    print('call %.3f ms | step: first start 0, last start %.3f, last finish %.3f | render: first start %.3f, last start %.3f, last end %.3f' % (
        e0.elapsed_time(e1), ms(s0.max()), ms(s1.max()), ms(r0.min()), ms(r0.max()), ms(r1.max())))
    dur = (s1 - s0) / 1e6
    print('  step duration per env: mean %.3f p50 %.3f p90 %.3f p99 %.3f max %.3f ms; sum / 592 slots = %.3f ms' % (
        dur.mean(), *np.percentile(dur, [50, 90, 99]), dur.max(), dur.sum() / 592))
    rd = (r1 - r0) / 1e6
    print('  render duration per env: mean %.3f p50 %.3f p99 %.3f max %.3f ms; wait finish -> render start: mean %.3f p50 %.3f p99 %.3f max %.3f ms' % (
        rd.mean(), np.percentile(rd, 50), np.percentile(rd, 99), rd.max(), ((r0 - s1) / 1e6).mean(),
        *np.percentile((r0 - s1) / 1e6, [50, 99]), ((r0 - s1) / 1e6).max()))
    edges = np.arange(0, ms(r1.max()) + 0.25, 0.25)
    fin = np.histogram(ms(s1), edges)[0]
    ren = np.histogram(ms(r1), edges)[0]
    run = [(int(((ms(s0) <= a) & (ms(s1) > a)).sum())) for a in edges[:-1]]
    print('  t(ms)   steps running  steps finished  frames drawn')
    for a, x, y, z in zip(edges[:-1], run, fin, ren):
        print('  %5.2f   %6d  %6d  %6d' % (a, x, y, z))
    late = np.argsort(-s1)[:6]
    for i in late:
        print('  env %5d  step %.3f -> %.3f  render %.3f -> %.3f' % (i, ms(s0[i]), ms(s1[i]), ms(r0[i]), ms(r1[i])))
