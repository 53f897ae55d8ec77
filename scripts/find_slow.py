"""Replays bench.py's sequence and reports the slowest envs (diagnostic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from moog_b200.batched_env import BatchedEnvironment
scene='falling_balls20'; E=4096
cfg = bench._scene_config(scene)
states = bench._host_states(cfg, 128, 1234)
env = BatchedEnvironment(**cfg, num_envs=E, device='cuda:0', seed=1234, initial_states=states)
env.reset(); eng=env.engine
g = torch.Generator(device='cpu').manual_seed(1234)
act = torch.randint(0, 5, (E, env.action_dim), generator=g).to(torch.float64).to('cuda:0')
for t in range(8):
    before = {k: getattr(eng.state,k).clone() for k in eng.state.KEYS}
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); eng.env_step(act, want_counters=True); e1.record(); torch.cuda.synchronize()
    c = eng.counters.cpu().numpy()
    cyc = c[:,4]
    top = np.argsort(-cyc)[:4]
    print('step %d %.2f ms  cycles median %.3g max %.3g  top envs %s calls %s coll %s' % (t, e0.elapsed_time(e1), np.median(cyc), cyc.max(), top.tolist(), c[top,0].tolist(), c[top,2].tolist()), flush=True)
    if cyc.max() > 50*np.median(cyc):
        i = int(top[0])
        np.savez('gpurun_out/slow_env.npz', blob=np.frombuffer(env.program.blob, dtype=np.uint8), **{k: v[i].cpu().numpy() for k,v in before.items()})
        d = before['dyn'][i].cpu().numpy()
        print('env', i, 'dyn before:\n', np.array2string(d, precision=4, max_line_width=200))
        break
