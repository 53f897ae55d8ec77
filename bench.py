#!/usr/bin/env python
"""bench.py -- batched env-steps/s including the 64x64 RGB render (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--scene falling_balls20|synthetic32] [--envs E]

A "step" is one Environment.step of every env of the batch (physics K substeps
+ rules + task) followed by the PILRenderer frame of every env.  N=1 workload:
BASELINE.json configs[1], `falling_balls20` with 4096 envs per GPU (weak
scaling: every rank owns its own 4096 envs; no data-path collective, one NCCL
all_reduce of the 4 episode statistics at the end of the timed region).

Prints ONE JSON line (see the contract in the task statement).  `--impl
reference` times the CPU oracle port of the same path on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'batched env-steps/sec incl. 64x64 RGB render'
UNIT = 'env-steps/s'

# BASELINE.json configs as bench scenes (moog_b200/configs/<scene>.py): envs per GPU, the episode
# length used to stagger the phases, untimed burn-in steps, host-generated initial states.
# The headline (and the default) is configs[1], falling_balls20 at 4096 envs; the others are
# reported through the same code path with `--scene` (profiles/r02_other_scenes.jsonl).
SCENES = {
    'falling_balls20': dict(envs=4096, episode=100, burn_in=130, pool=512),
    'colliding_predators84': dict(envs=16384, episode=200, burn_in=60, pool=512),
    'cleanup64': dict(envs=8192, episode=200, burn_in=60, pool=256),
    'pacman64': dict(envs=8192, episode=200, burn_in=60, pool=256),
    'synthetic32': dict(envs=131072, episode=200, burn_in=60, pool=512),   # 2^20 envs over 8 GPUs
}


def _metric(prog):
    if prog.render is None:
        return 'batched env-steps/sec, state only (no observer)'
    return 'batched env-steps/sec incl. {}x{} RGB render'.format(prog.render['height'], prog.render['width'])


def _random_actions(prog, n, rng):
    """[n, action_dim] float64: a random action per action-space component, as the reference's
    `random_action()` draws them (Grid: one of 5 indices; Joystick: U[-1, 1]^2; SetPosition: U[0, 1]^2)."""
    out = np.zeros((n, max(prog.action_dim, 1)))
    for _, kind, off, width in getattr(prog, 'action_layout', []):
        if kind == 'Grid':
            out[:, off] = rng.randint(0, 5, size=n)
        elif kind == 'SetPosition':
            out[:, off:off + width] = rng.uniform(0, 1, size=(n, width))
        else:
            out[:, off:off + width] = rng.uniform(-1, 1, size=(n, width))
    return out


def _physical_cores():
    try:
        pairs = set()
        phys = core = None
        for line in open('/proc/cpuinfo'):
            if line.startswith('physical id'):
                phys = line.split(':')[1].strip()
            elif line.startswith('core id'):
                core = line.split(':')[1].strip()
            elif not line.strip():
                if phys is not None and core is not None:
                    pairs.add((phys, core))
                phys = core = None
        return len(pairs) or None
    except Exception:  # pylint: disable=broad-except
        return None

# algorithmic bytes per env-step
#   B = S*(32 R + 32 W dynamic + 32 R static) + 8 (action) + 12 (reward, step_type, discount) + H*W*3
ALGO_BYTES = {
    'falling_balls20': dict(state=24 * 96 + 20, frame=64 * 64 * 3),
    'synthetic32': dict(state=32 * 96 + 20, frame=64 * 64 * 3),
}


def _scene_config(scene, state_only=False):
    import moog_b200  # noqa: F401  (puts the MOOG-compatible `moog` package on sys.path)
    import importlib
    mod = importlib.import_module('moog_b200.configs.' + scene)
    cfg = mod.get_config()
    if state_only:
        cfg['observers'] = {}
    return cfg


def _host_states(config, n, seed):
    np.random.seed(seed)
    return [config['state_initializer']() for _ in range(n)]


_REAL_STDOUT = None


def _quiet_stdout():
    """Everything written to fd 1 from here on (NCCL's own "NCCL version ..." banner, library
    chatter) goes to stderr; the one JSON line is written to the real stdout by _emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    text = json.dumps(line, default=_json_default) + '\n'
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, text.encode())


def _json_default(o):
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    raise TypeError(type(o).__name__)


def _ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the
    committed `ncu --set full` summary of this round (profiles/), or None."""
    import glob
    found = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r[0-9][0-9]_{}_ncu.json'.format(kernel))))
    path = found[-1] if found else ''          # the latest round's capture
    try:
        with open(path) as f:
            d = json.load(f)
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        total = 0.0
        for key in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            val, unit = d[key].split()
            total += float(val) * scale[unit]
        return total
    except Exception:  # pylint: disable=broad-except
        return None


def _peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), 'measured'
    except Exception:  # pylint: disable=broad-except
        return 6650.0, 'fallback'


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.first = 0

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '20'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:  # pylint: disable=broad-except
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Samples from here on (and the one just before, taken under the same load) count."""
        self.first = max(len(self.lines) - 1, 0)

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # pylint: disable=broad-except
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines[self.first:]:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None,
                    sm_max_mhz=float(np.max(mx)) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ---------------------------------------------------------------------------

class CpuArm(object):
    """The oracle port (oracle/moog_oracle.c + pil_oracle.c) on the host cores.

    One Oracle of `envs_per_thread` envs per thread (ctypes releases the GIL).
    Thread t is pre-advanced by t/threads of an episode (untimed), and envs are
    re-initialised when their episode ends, so any timed window sees the same
    uniform mix of episode phases as the GPU arm.
    """

    def __init__(self, scene, envs_per_thread=8, threads=None, episode=100, seed=1234, state_only=False):
        from moog_b200 import compiler
        from oracle.oracle import Oracle, lib as orc_lib
        orc_lib()
        self.threads = threads or (os.cpu_count() or 1)
        self.n = envs_per_thread
        self.episode = episode
        config = _scene_config(scene, state_only)
        states = _host_states(config, max(16, envs_per_thread), seed)
        self.prog = compiler.compile_config(config, states)
        base = compiler.pack_states(self.prog, states)
        self.keys = ('dyn', 'stat', 'meta', 'cnt', 'envi', 'envf', 'vtx')
        self.init = []
        self.oracles = []
        for t in range(self.threads):
            idx = (np.arange(self.n) + t * self.n) % len(states)
            arr = {k: np.ascontiguousarray(base[k][idx]) for k in self.keys}
            self.init.append(arr)
            o = Oracle(self.prog, arr)
            o.post_reset()
            self.oracles.append(o)
        self.actions = _random_actions(self.prog, self.n, np.random.RandomState(seed))
        self.render = self.prog.render is not None
        self._pool = None

    def _advance(self, t, steps, render=True):
        o = self.oracles[t]
        p = self.prog
        rng = np.random.RandomState(977 + t)
        for _ in range(steps):
            # the uniforms behind RandomForce / RandomMazeWalk / ModifySprites(sample_one)
            noise = rng.uniform(size=(self.n, p.K, p.noise_dim)) if p.noise_dim else None
            rule_noise = rng.uniform(size=(self.n, p.rule_noise_dim)) if p.rule_noise_dim else None
            _, st = o.step(self.actions, noise=noise, rule_noise=rule_noise)
            if render and self.render:
                o.render()
            if (st == 2).all():      # environment.py:100-101: next step() is reset()
                for k in self.keys:
                    getattr(o, k)[...] = self.init[t][k]
                o.post_reset()
                if render and self.render:
                    o.render()
        return True

    def _map(self, fn):
        from concurrent.futures import ThreadPoolExecutor
        if self._pool is None:
            self._pool = ThreadPoolExecutor(max_workers=self.threads)
        return list(self._pool.map(fn, range(self.threads)))

    def stagger(self):
        self._map(lambda t: self._advance(t, (t * self.episode) // self.threads, render=False))

    def run(self, steps):
        """Advances every env by `steps` env-steps; returns (env-steps/s, seconds, env-steps)."""
        t0 = time.perf_counter()
        self._map(lambda t: self._advance(t, steps))
        dt = time.perf_counter() - t0
        total = self.threads * self.n * steps
        return total / dt, dt, total


def _apply_scene_defaults(args):
    d = SCENES.get(args.scene, SCENES['falling_balls20'])
    if args.envs is None:
        args.envs = d['envs']
    if args.episode is None:
        args.episode = d['episode']
    if args.burn_in is None:
        args.burn_in = d['burn_in']
    if args.pool is None:
        args.pool = d['pool']


def _reference_python():
    """The UNMODIFIED reference timed in the build container (scripts/time_reference_python.py): it
    does not travel to the GPU box, so this committed measurement is quoted beside the port."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r02_reference_python.json')) as f:
            d = json.load(f)
        pool = d['pool_falling_balls20']
        return {'value': pool['env_steps_per_s'], 'unit': UNIT, 'cores': pool['workers'],
                'per_core': pool['env_steps_per_s_per_core'],
                'where': 'build container ({} cores), not this box: /root/reference is not present here'.format(d['cores']),
                'what': pool['what'], 'note': d['note']}
    except Exception:  # pylint: disable=broad-except
        return None


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    arm = CpuArm(args.scene, envs_per_thread=8, episode=args.episode, state_only=args.state_only)
    arm.stagger()
    per_step = 5   # env-steps per env per timed step: 20 steps cover a whole episode
    for _ in range(args.warmup):
        arm.run(1)
    vals = []
    t0 = time.perf_counter()
    total = 0
    for _ in range(args.steps):
        v, dt, n = arm.run(per_step)
        vals.append(v)
        total += n
    wall = time.perf_counter() - t0
    value = total / wall
    sample = ('{} threads x {} envs x {} env-steps per timed step of {} (oracle/moog_oracle.c + pil_oracle.c, '
              'uniform episode-phase mix)').format(arm.threads, arm.n, per_step, args.scene)
    r = arm.prog.render
    line = {
        'impl': 'reference', 'metric': _metric(arm.prog), 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * wall / max(args.steps, 1), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': args.scene, 'envs_per_gpu': args.envs, 'sprites': arm.prog.n_slots,
                   'substeps': arm.prog.K,
                   'image': '{}x{}x3'.format(r['height'], r['width']) if r else 'none (state only)',
                   'parallelism': 'env-sharded x{}'.format(args.gpus),
                   'note': 'the reference arm: the CPU port of Environment.step + PILRenderer '
                           '(oracle/moog_oracle.c + pil_oracle.c) on all host threads, each timed step a bounded '
                           'sample of this workload (see cpu_baseline.sample)'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': arm.threads, 'physical_cores': _physical_cores(),
                         'kind': 'port', 'sample': sample,
                         **({'reference_python': _reference_python()} if args.scene == 'falling_balls20' and _reference_python() else {})},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    _emit(line)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist
    from moog_b200 import capi
    from moog_b200 import dist as mdist
    from moog_b200.batched_env import BatchedEnvironment

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    affinity = _pin_to_gpu_numa_node(local_rank)     # before any pinned allocation (first touch)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    config = _scene_config(args.scene, args.state_only)
    E = args.envs
    seed = mdist.rank_seed(1234, rank)
    states = _host_states(config, args.pool, seed)
    env = BatchedEnvironment(**config, num_envs=E, device=dev, seed=seed, initial_states=states,
                             reset_mode=args.reset_mode)
    eng = env.engine
    prog = env.program
    ad = env.action_dim
    has_frames = prog.render is not None
    H, W = (prog.render['height'], prog.render['width']) if has_frames else (0, 0)

    g = torch.Generator(device='cpu').manual_seed(seed)
    host_actions = torch.from_numpy(_random_actions(prog, E, np.random.RandomState(seed & 0x7fffffff))).pin_memory()
    dev_actions = host_actions.to(dev)
    host_frames = torch.empty((E, H, W, 3), dtype=torch.uint8).pin_memory() if has_frames else None
    host_reward = torch.empty(E, dtype=torch.float32).pin_memory()
    host_step_type = torch.empty(E, dtype=torch.int32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    env.reset()
    # Burn-in: every env is force-reset once at a random time of the first
    # `episode` steps (host sets its reset_next word), so that by the timed region
    # the batch holds a uniform mix of episode phases -- free fall, first contacts,
    # settled piles, and ~1/episode of the envs resetting per step -- instead of
    # N copies of step 0.
    episode = args.episode
    phase = torch.randint(0, episode, (E,), generator=g).to(dev)
    for t in range(args.burn_in):
        if t < episode:
            eng.state.envi[:, 1] = torch.where(phase == t, torch.ones_like(phase, dtype=torch.int32),
                                               eng.state.envi[:, 1])
        eng.env_step(dev_actions, sample_resets=(args.reset_mode == 'device'))
    eng.stats.zero_()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- device-resident arm ------------------------------------------------
    # --render auto: the step call is asked for the frames (moog_step_io.frames) and the library
    # draws them inside the step kernel when a canvas fits next to the env record at no cost in
    # residency, else with the render kernel; --render separate: step call, then render call
    fused = has_frames and args.render == 'auto' and eng.dev_program.step_draws_frames(E)

    def device_step():
        if args.render == 'auto' and has_frames:
            eng.env_step(dev_actions, sample_resets=(args.reset_mode == 'device'), frames=True)
        else:
            eng.env_step(dev_actions, sample_resets=(args.reset_mode == 'device'))

    def device_render():
        if args.render != 'auto' and has_frames:
            eng.render()

    # nvidia-smi needs a moment to deliver its first line: the sampler starts before the warm-up,
    # the warm-up goes on (bounded) until a sample taken under this load has arrived, and the
    # samples from that one on are the ones reported
    sampler = ClockSampler(local_rank)
    if not args.no_clocks:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        device_step()
        device_render()
    torch.cuda.synchronize()
    if not args.no_clocks and sampler.proc is not None:
        t_wait = time.perf_counter()
        n_before = len(sampler.lines)
        while len(sampler.lines) < n_before + 2 and time.perf_counter() - t_wait < 3.0:
            device_step()
            device_render()
            torch.cuda.synchronize()
        sampler.mark()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    launches0 = capi.launch_count()
    barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 255)          # evict L2 between timed iterations (untimed)
        ev[k][0].record()
        device_step()
        ev[k][1].record()
        device_render()
        ev[k][2].record()
    stats = env.episode_stats(reduce=True)   # the only collective: 4 doubles
    torch.cuda.synchronize()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = capi.launch_count() - launches0
    clocks = sampler.stop()
    ms_step_k = [ev[k][0].elapsed_time(ev[k][1]) for k in range(args.steps)]
    ms_rend_k = [ev[k][1].elapsed_time(ev[k][2]) for k in range(args.steps)]
    ms_total = float(np.sum(ms_step_k) + np.sum(ms_rend_k))
    if args.verbose:
        sys.stderr.write('step ms %s\nrender ms %s\n' % (ms_step_k, ms_rend_k))
    ms_total_max = mdist.max_over_ranks(ms_total, dev)
    value = world * E * args.steps / (ms_total_max * 1e-3)

    # ---- the two kernels on their own (roofline legs) ----------------------------
    # In the timed region above the render kernel runs behind the step kernel and overlaps its
    # tail, so one interval covers both.  For the per-kernel durations the same batch is stepped
    # without frames (moog_step_kernel alone) and rendered on its own (moog_render_kernel), L2
    # flushed before each; these steps are not part of `value`.
    kev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    for k in range(args.steps):
        flush.fill_(k & 255)
        kev[k][0].record()
        eng.env_step(dev_actions, sample_resets=(args.reset_mode == 'device'))
        kev[k][1].record()
        if has_frames:
            eng.render()
        kev[k][2].record()
    torch.cuda.synchronize()
    step_only_ms = float(np.mean([kev[k][0].elapsed_time(kev[k][1]) for k in range(args.steps)]))
    render_only_ms = float(np.mean([kev[k][1].elapsed_time(kev[k][2]) for k in range(args.steps)]))

    # ---- end-to-end arm: host actions in, host TimeStep (frames included) out --
    # (a) one step per call (`step_to_host`: the host is handed step k before step k+1 is enqueued)
    from moog_b200.batched_env import TimeStep
    host_ts = TimeStep(host_step_type, host_reward, None, {'image': host_frames} if has_frames else {})
    for _ in range(10):   # warm-up; with --e2e-frames auto these calls also pick the frame path
        env.step_to_host(host_actions, host_ts, chunks=args.e2e_chunks, frames=args.e2e_frames)
    torch.cuda.synchronize()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        # H2D of the actions, the step, the frames into the pinned host image (written by the
        # kernels themselves over PCIe as the envs finish -- 'mapped' -- or drawn in HBM and
        # copied); returns when the host buffers hold the whole TimeStep (the caller owns it)
        env.step_to_host(host_actions, host_ts, chunks=args.e2e_chunks, frames=args.e2e_frames)
    e1.record()
    torch.cuda.synchronize()
    e2e_sync_value = world * E * args.steps / (mdist.max_over_ranks(e0.elapsed_time(e1), dev) * 1e-3)
    sync_frames = (env._auto_choice or 'device') if args.e2e_frames == 'auto' else args.e2e_frames
    # (b) the public pipelined API (`env.host_pipeline(depth)`): the same per-step traffic -- that
    # step's actions host -> device from pinned memory, its whole TimeStep device -> host into
    # pinned memory -- with step k+1 enqueued before the host collects step k.  Both frame
    # transports are timed over a few steps on this box and the faster one is measured.
    def pipelined(frames_mode, steps, timed):
        pipe = env.host_pipeline(depth=args.e2e_depth, frames=frames_mode)
        sink = 0
        torch.cuda.synchronize()
        if timed:
            barrier()
        p0 = torch.cuda.Event(enable_timing=True)
        p1 = torch.cuda.Event(enable_timing=True)
        p0.record()
        for k in range(steps):
            if k >= args.e2e_depth:
                ts = pipe.collect()
                sink += int(ts.step_type[0])          # the host reads the result
            pipe.submit(host_actions)
        for _ in range(min(args.e2e_depth, steps)):
            ts = pipe.collect()
            sink += int(ts.step_type[0])
        main = torch.cuda.current_stream(dev)
        main.wait_event(pipe.last_event())            # (the copy stream's tail, for frames='device')
        p1.record()
        torch.cuda.synchronize()
        return p0.elapsed_time(p1), ts
    trial = {}
    modes = ('mapped', 'device') if args.e2e_frames in ('auto',) else (
        ('mapped',) if args.e2e_frames == 'mapped' else ('device',))
    for m in modes:
        pipelined(m, 6, False)                         # warm-up (allocates the pinned slots)
        trial[m] = pipelined(m, 8, False)[0]
    pipe_mode = min(trial, key=trial.get)
    pipe_ms, last_ts = pipelined(pipe_mode, args.steps, True)
    e2e_value = world * E * args.steps / (mdist.max_over_ranks(pipe_ms, dev) * 1e-3)
    h2d = host_actions.numel() * host_actions.element_size()
    d2h = ((last_ts.observation['image'].numel() if has_frames else 0) + last_ts.reward.numel() * 4
           + last_ts.step_type.numel() * 4 + last_ts.discount.numel() * 4)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = _peaks()
    ab = ALGO_BYTES.get(args.scene, dict(state=prog.n_slots * 96 + 20, frame=H * W * 3))
    call_ms = float(np.mean(ms_step_k))     # the step call of the timed region (with the frames under --render auto)
    step_ms = step_only_ms                  # moog_step_kernel (+ the 11 us order kernel) on its own
    rend_ms = render_only_ms                # moog_render_kernel on its own
    # algorithmic bytes of one step-kernel launch: the record in and out
    step_bytes = ab['state']
    step_gbs = E * step_bytes / (step_ms * 1e-3) / 1e9
    rend_gbs = E * (ab['frame'] + prog.n_slots * 32) / (rend_ms * 1e-3) / 1e9 if has_frames else 0.0
    record_bytes = eng.state.nbytes() // E
    line = {
        'metric': _metric(prog), 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_total_max / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': args.scene, 'envs_per_gpu': E, 'sprites': prog.n_slots,
                   'substeps': prog.K, 'image': '{}x{}x3'.format(H, W) if has_frames else 'none (state only)',
                   'parallelism': 'env-sharded x{}'.format(world),
                   'l2': 'flushed between timed iterations (256 MiB fill, untimed)',
                   'phases': 'uniform mix of episode phases after {} burn-in steps with staggered resets'.format(args.burn_in),
                   'state_record_bytes': int(record_bytes),
                   'resets': ('device-side sampler' if args.reset_mode == 'device' else
                              'pool of {} host-generated initial states'.format(len(states))),
                   'step_launch': dict(zip(('resident_envs_per_sm', 'warps_per_env', 'smem_bytes_per_env'),
                                           eng.dev_program.step_launch_info(E))),
                   'frames': ('separate render call after the step call' if args.render != 'auto' else
                              'drawn inside moog_step_kernel' if fused else
                              'render kernel launched behind the step kernel (programmatic stream serialization), '
                              'drawing the envs in finishing order while the longest envs are still stepped'),
                   'e2e': 'env.host_pipeline(depth={}): step k+1 is enqueued before the host collects the TimeStep of '
                          'step k (actions one observation old); frames {}'.format(
                              args.e2e_depth, 'stored by the kernels straight into the pinned host image'
                              if pipe_mode == 'mapped' else 'drawn in HBM, copied by a copy engine on a second stream'),
                   'e2e_frames': pipe_mode, 'e2e_pipeline_depth': args.e2e_depth,
                   'e2e_frames_trial_ms': {k: round(v / 8, 4) for k, v in trial.items()},
                   'cpu_affinity': affinity},
        'roofline': {'bound': 'hbm', 'kernel': 'moog_step_kernel', 'achieved': step_gbs, 'peak': peak,
                     'unit': 'GB/s', 'frac': step_gbs / peak,
                     'traffic': (_ncu_traffic('step_kernel')
                                 if args.scene == 'falling_balls20' and E == 4096 else None),
                     'traffic_note': ('bytes per launch, ncu --set full of this workload (profiles/): one read of the '
                                      '13.4 KB env records (9.9 KB of each is the cached world vertices, state that the '
                                      'algorithmic figure does not count) + the written-back part that left L2; no re-reads'),
                     'algorithmic_bytes_per_launch': E * step_bytes,
                     'peak_source': peak_src,
                     'algorithmic_bytes_per_env_step': step_bytes, 'kernel_ms': step_ms,
                     'kernel_ms_note': 'each kernel timed on its own after the timed region (CUDA events, L2 flushed)',
                     'render_kernel': {'kernel': 'moog_render_kernel', 'achieved': rend_gbs, 'frac': rend_gbs / peak,
                                       'algorithmic_bytes_per_env_step': ab['frame'] + prog.n_slots * 32,
                                       'traffic': (_ncu_traffic('render_kernel')
                                                   if args.scene == 'falling_balls20' and E == 4096 else None),
                                       'kernel_ms': rend_ms},
                     'share_of_step': {'moog_step_kernel': step_ms / (step_ms + rend_ms),
                                       'moog_render_kernel': rend_ms / (step_ms + rend_ms)},
                     'whole_step': {'algorithmic_bytes_per_env_step': step_bytes + (ab['frame'] if has_frames else 0),
                                    'achieved': E * (step_bytes + (ab['frame'] if has_frames else 0)) /
                                                (ms_total_max / args.steps * 1e-3) / 1e9,
                                    'frac': E * (step_bytes + (ab['frame'] if has_frames else 0)) /
                                            (ms_total_max / args.steps * 1e-3) / 1e9 / peak},
                     'timed_step_call_ms': call_ms,
                     'overlap_ms': step_ms + rend_ms - ms_total_max / args.steps},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h)},
        'e2e_one_step_per_call': {'value': e2e_sync_value, 'unit': UNIT, 'api': 'BatchedEnvironment.step_to_host',
                                  'frames': sync_frames},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'episode_stats': [float(x) for x in stats.tolist()],
        'wall_s_timed_region': t_wall,
    }
    if not args.no_cpu:
        arm = CpuArm(args.scene, envs_per_thread=8, episode=args.episode, state_only=args.state_only)
        arm.stagger()
        # a bounded sample: whole episodes for the headline scene, ~20 s of work for the others
        cpu_steps = args.episode * args.cpu_episodes if args.scene == 'falling_balls20' else max(args.episode // 2, 20)
        v, dt, n = arm.run(cpu_steps)
        line['cpu_baseline'] = {
            'value': v, 'unit': UNIT, 'cores': arm.threads, 'physical_cores': _physical_cores(), 'kind': 'port',
            'sample': '{} env-steps of {}{}: {} threads x 8 envs x {} env-steps each '
                      '(uniform phase mix), {:.1f} s'.format(n, args.scene, ' incl. render' if has_frames else ', state only',
                                                             arm.threads, cpu_steps, dt)}
        ref_py = _reference_python() if args.scene == 'falling_balls20' else None
        if ref_py:
            line['cpu_baseline']['reference_python'] = ref_py
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def _pin_to_gpu_numa_node(local_rank):
    """Bind this rank's host threads to the CPUs NVML lists as local to its GPU, so that the pinned
    TimeStep buffers (first touch) and the submitting thread sit on the GPU's NUMA node.  Returns a
    short description for the JSON line; any failure leaves the affinity alone."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + uuid).encode())
        except Exception:  # pylint: disable=broad-except
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed and len(allowed) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            return 'bound to the {} CPUs local to GPU {} ({}..{})'.format(len(allowed), local_rank, allowed[0], allowed[-1])
        return 'unchanged ({} CPUs, all local to GPU {})'.format(len(os.sched_getaffinity(0)), local_rank)
    except Exception as exc:  # pylint: disable=broad-except
        return 'unchanged ({})'.format(type(exc).__name__)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--scene', default='falling_balls20')
    ap.add_argument('--envs', type=int, default=None, help='envs per GPU (default: the scene\'s BASELINE.json size)')
    ap.add_argument('--pool', type=int, default=None, help='host-generated initial states')
    ap.add_argument('--episode', type=int, default=None, help='episode length used to stagger phases')
    ap.add_argument('--burn-in', type=int, default=None, help='untimed steps before the timed region')
    ap.add_argument('--state-only', action='store_true', help='observers = {}: no frames (BASELINE configs[4])')
    ap.add_argument('--no-clocks', action='store_true')
    ap.add_argument('--reset-mode', default='pool', choices=['pool', 'device'],
                    help="'device': resetting envs draw their generated sprites on the GPU")
    ap.add_argument('--e2e-chunks', type=int, default=4,
                    help="env ranges the e2e arm renders / copies in (--e2e-frames chunked)")
    ap.add_argument('--render', default='auto', choices=['auto', 'separate'],
                    help="'auto': the step call draws the frames (inside the step kernel when they fit); "
                         "'separate': step call, then render call")
    ap.add_argument('--e2e-frames', default='auto', choices=['auto', 'mapped', 'device', 'chunked'],
                    help='how the e2e arm brings the frames to the host (BatchedEnvironment.step_to_host)')
    ap.add_argument('--e2e-depth', type=int, default=2, help='steps in flight in the end-to-end arm (HostPipeline)')
    ap.add_argument('--verbose', action='store_true')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--cpu-episodes', type=int, default=6,
                    help='whole episodes every env of the cpu_baseline leg walks (6: about 10 s on 16 threads)')
    args = ap.parse_args()
    _apply_scene_defaults(args)
    _quiet_stdout()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
