"""The shipped first_person_predators_prey program (from its golden) at 4096 envs: env-steps/s with frames."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import moog_b200
from moog_b200.batched_env import Engine
from tests import util
from tests.test_create_sprites import _load_spawn
g = _load_spawn('first_person'); prog = g['program']
N = 4096
arrays = util.tile_state(util.state_at(g, 0, prefix='init'), N)
eng = Engine(prog, N, 'cuda:0', seed=3)
eng.state.upload(arrays); eng.set_pool({k: arrays[k][:2] for k in util.STATE_KEYS})
eng.post_reset()
act = torch.rand((N, 2), dtype=torch.float64, device='cuda:0') * 2 - 1
for t in range(60):
    eng.env_step(act, auto_reset=True, frames=True)
torch.cuda.synchronize()
t0 = time.time()
for t in range(40):
    eng.env_step(act, auto_reset=True, frames=True)
torch.cuda.synchronize()
dt = (time.time() - t0) / 40
st = eng.state.download()
print('first_person 4096 envs: %.2f ms/step  %.0f env-steps/s' % (dt * 1e3, N / dt), 'cnt max', st['cnt'].max(axis=0)[:5], 'err envs', int((st['envi'][:, 2] != 0).sum()), 'episodes', st['envi'][:, 3].max())
import ctypes
from moog_b200 import capi
a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
capi.lib().moog_step_launch_info(eng.dev_program.handle, N, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
print('resident/SM', a.value, 'warps/env', b.value, 'smem/env', c.value)
