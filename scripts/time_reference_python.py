"""Times the UNMODIFIED reference (pure Python MOOG through oracle/shims: matplotlib / dm_env
stand-ins) in THIS container -- the reference does not travel to the GPU box, so this is the only
place it can be timed.  Writes profiles/r02_reference_python.json,
which bench.py quotes as `cpu_baseline.reference_python`.

  (i)  tests/runtime_benchmark.py verbatim on the shipped pong config (one core);
  (ii) a multiprocessing pool, one reference `Environment` of the bench scene per worker, all host
       cores, random actions, 2 warm-up resets, whole episodes with the 64x64 PILRenderer frame.
"""
import json
import multiprocessing as mp
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
SHIMS = os.path.join(ROOT, 'oracle', 'shims')


def _worker(arg):
    scene, steps, seed = arg
    sys.path.insert(0, ROOT)
    from oracle import refenv
    refenv.activate()
    import importlib
    import numpy as np
    import moog_b200  # noqa: F401  (configs package)
    from moog import environment
    np.random.seed(seed)
    cfg = importlib.import_module('moog_b200.configs.' + scene).get_config()
    env = environment.Environment(**cfg)
    env.reset()
    env.reset()
    t0 = time.perf_counter()
    n = 0
    while n < steps:
        ts = env.step(env.action_space.random_action())
        n += 1
    return n, time.perf_counter() - t0


def main():
    out = {'host': 'build container', 'cores': os.cpu_count(),
           'note': 'unmodified /root/reference through oracle/shims (pure-Python matplotlib Path / Affine2D '
                   'stand-ins, real Pillow): pessimistic for the geometry, genuine for the renderer'}
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', PYTHONPATH=os.pathsep.join([SHIMS, REF]))
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, 'runtime_benchmark.py', '--config=moog_demos.example_configs.pong'],
                       cwd=os.path.join(REF, 'tests'), env=env, capture_output=True, text=True, timeout=3600)
    text = r.stdout if r.returncode == 0 else r.stdout + r.stderr      # (stderr holds tqdm's progress bars)
    out['runtime_benchmark_pong'] = {'seconds': time.perf_counter() - t0, 'returncode': r.returncode,
                                     'output': [l for l in text.splitlines() if l.strip() and '%|' not in l][-60:]}
    scene, steps = 'falling_balls20', int(os.environ.get('REF_STEPS', '100'))
    workers = os.cpu_count() or 1
    t0 = time.perf_counter()
    with mp.get_context('spawn').Pool(workers) as pool:
        res = pool.map(_worker, [(scene, steps, 1234 + 1000 * w) for w in range(workers)])
    wall = time.perf_counter() - t0
    total = sum(n for n, _ in res)
    out['pool_' + scene] = {'workers': workers, 'env_steps': total, 'wall_s': wall,
                            'env_steps_per_s': total / max(dt for _, dt in res),
                            'env_steps_per_s_per_core': total / sum(dt for _, dt in res),
                            'what': '{} workers x 1 reference Environment x {} env-steps incl. 64x64 PILRenderer frame, '
                                    'random actions, auto-reset'.format(workers, steps)}
    path = os.path.join(ROOT, 'profiles', 'r02_reference_python.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != 'runtime_benchmark_pong'}, indent=1))
    print('\n'.join(out['runtime_benchmark_pong']['output'][-25:]))


if __name__ == '__main__':
    main()
