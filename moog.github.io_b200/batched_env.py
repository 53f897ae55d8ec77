"""BatchedEnvironment: N independent MOOG environments stepped on one B200.

Mirrors `moog.environment.Environment` (reference moog/environment.py:28-131):
same constructor arguments -- so `BatchedEnvironment(**get_config(level),
num_envs=N)` works on an unchanged MOOG config -- and the same `reset()` /
`step(action)` / `observation()` protocol, with every TimeStep field batched:

    step_type   int32  [N]   0 FIRST, 1 MID, 2 LAST        (dm_env.StepType)
    reward      float32[N]   NaN on FIRST steps            (dm_env: None)
    discount    float32[N]   1 mid, 0 last, NaN first
    observation {'image': uint8 [N, H, W, 3]}              (PILRenderer)

An env whose previous step returned LAST is reset by the next `step()` call and
ignores that call's action (environment.py:100-101).  Initial states come from
a pool produced by the config's own `state_initializer` on the host and kept
on the device; which pool entry a resetting env receives is drawn on the device.

PyTorch is used for device memory and streams only.  All arithmetic happens in
libmoog_b200.so (include/moog_b200.h); there is no fallback path.
"""
import collections
import contextlib
import ctypes
import time

import numpy as np
import torch

from . import capi
from . import compiler

TimeStep = collections.namedtuple(
    'TimeStep', ['step_type', 'reward', 'discount', 'observation'])

STEP_FIRST, STEP_MID, STEP_LAST = 0, 1, 2

_ERR_TEXT = {
    1: 'ValueError: collision normal is not a unit vector (collisions.py:323-326)',
    2: 'RuntimeError: _position_correction would not terminate (collisions.py:740)',
    4: 'ValueError: TetherZippedLayers layers differ in length (tether_physics.py:192-196)',
    8: 'RuntimeError: layer capacity exceeded',
    16: 'ValueError: Object is not on the maze grid (maze_physics.py:93-104)',
    32: 'RecursionError: max_recursion_depth exceeded trying to initialize a non-overlapping sprite '
        '(sprite_generators.py:92-98) / ValueError: maximum number of tries exceeded (distributions.py:341-349)',
    64: 'ValueError: There must be an even number of portals (portal.py:49-52)',
    128: 'IndexError: a condition / reward function indexed state[layer][i] beyond the sprites the layer holds',
}

_STATE_DTYPES = dict(dyn=torch.float64, stat=torch.float64, meta=torch.int32,
                     cnt=torch.int32, envi=torch.int32, envf=torch.float64,
                     vtx=torch.float64)


class DeviceState(object):
    """The SoA state record of n envs as torch tensors on one device."""

    KEYS = ('dyn', 'stat', 'meta', 'cnt', 'envi', 'envf', 'vtx')

    def __init__(self, program, n, device):
        S, VT = program.n_slots, max(program.n_vtx, 1)
        self.n = n
        self.device = torch.device(device)
        z = lambda shape, k: torch.zeros(shape, dtype=_STATE_DTYPES[k], device=self.device)
        self.dyn = z((n, compiler.DYN_FIELDS, S), 'dyn')
        self.stat = z((n, compiler.STAT_FIELDS, S), 'stat')
        self.meta = z((n, compiler.META_FIELDS, S), 'meta')
        self.cnt = z((n, compiler.MAX_LAYERS), 'cnt')
        self.envi = z((n, compiler.ENVI_WORDS), 'envi')
        self.envf = z((n, max(program.n_envf, 1)), 'envf')
        self.vtx = z((n, VT, 2), 'vtx')

    def upload(self, arrays, rows=None):
        """arrays: dict from compiler.pack_states (numpy); rows: env indices."""
        for k in self.KEYS:
            src = torch.from_numpy(np.ascontiguousarray(arrays[k])).to(
                self.device, dtype=_STATE_DTYPES[k])
            dst = getattr(self, k)
            if rows is None:
                if src.shape != dst.shape:
                    raise ValueError('{}: shape {} != {}'.format(k, tuple(src.shape), tuple(dst.shape)))
                dst.copy_(src)
            else:
                dst[rows] = src

    def download(self):
        return {k: getattr(self, k).cpu().numpy() for k in self.KEYS}

    def clone(self):
        out = object.__new__(DeviceState)
        out.n, out.device = self.n, self.device
        for k in self.KEYS:
            setattr(out, k, getattr(self, k).clone())
        return out

    def copy_from(self, other):
        for k in self.KEYS:
            getattr(self, k).copy_(getattr(other, k))

    def struct(self, first=0):
        """`moog_state` of the envs [first, n): the row pointers of env `first`."""
        if first == 0:
            # the tensors are allocated once and only ever written in place
            key = tuple(getattr(self, k).data_ptr() for k in self.KEYS)
            cached = getattr(self, '_struct0', None)
            if cached is None or cached[0] != key:
                cached = (key, capi.MoogState(*key))
                self._struct0 = cached
            return cached[1]
        ptrs = []
        for k in self.KEYS:
            t = getattr(self, k)
            ptrs.append(t.data_ptr() + first * t.stride(0) * t.element_size())
        return capi.MoogState(*ptrs)

    def nbytes(self):
        return sum(getattr(self, k).numel() * getattr(self, k).element_size() for k in self.KEYS)


_NULL_CONTEXT = contextlib.nullcontext()


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class Engine(object):
    """Program + state + the raw C-ABI calls (used by BatchedEnvironment and
    directly by the parity tests / bench)."""

    def __init__(self, program, num_envs, device='cuda', seed=0):
        if not torch.cuda.is_available():
            raise capi.MoogError('moog_b200 needs a CUDA device; there is no CPU path')
        self.program = program
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise capi.MoogError('device must be a CUDA device')
        self.n = int(num_envs)
        self.seed = int(seed)
        with torch.cuda.device(self.device):
            self.dev_program = capi.DeviceProgram(program.blob)
        self.dev_program.set_option('seed', self.seed & 0x7fffffff)
        self.state = DeviceState(program, self.n, self.device)
        self.pool = None
        f32 = dict(dtype=torch.float32, device=self.device)
        self.reward = torch.zeros(self.n, **f32)
        self.discount = torch.zeros(self.n, **f32)
        self.step_type = torch.zeros(self.n, dtype=torch.int32, device=self.device)
        self.counters = torch.zeros((self.n, 8), dtype=torch.int64, device=self.device)
        self.stats = torch.zeros(4, dtype=torch.float64, device=self.device)
        self.frames = None
        if program.render is not None:
            r = program.render
            self.frames = torch.zeros((self.n, r['height'], r['width'], 3),
                                      dtype=torch.uint8, device=self.device)
        self._calls = 0
        self._pinned_ok = set()

    # -- helpers -----------------------------------------------------------
    def call_seed(self):
        """io.seed of the next env_step call (a different Philox key every call)."""
        return (self.seed * 0x9E3779B97F4A7C15 + self._calls) & 0xFFFFFFFFFFFFFFFF

    def _on_device(self):
        """Context that makes this engine's GPU current (free when it already is)."""
        index = self.device.index
        if index is None or torch.cuda.current_device() == index:
            return _NULL_CONTEXT
        return torch.cuda.device(self.device)

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_pool(self, arrays):
        n = arrays['dyn'].shape[0]
        self.pool = DeviceState(self.program, n, self.device)
        self.pool.upload(arrays)

    def _as_f64(self, x, width):
        if x is None:
            return None
        if not torch.is_tensor(x):
            x = torch.as_tensor(np.asarray(x))
        x = x.to(self.device, dtype=torch.float64, non_blocking=True).reshape(self.n, -1)
        if x.shape[1] != width:
            raise ValueError('expected [{}, {}], got {}'.format(self.n, width, tuple(x.shape)))
        return x.contiguous()

    # -- C-ABI calls ---------------------------------------------------------
    def post_reset(self, rule_noise=None):
        rn = self._as_f64(rule_noise, self.program.rule_noise_dim) if rule_noise is not None else None
        st = self.state.struct()
        with torch.cuda.device(self.device):
            capi.check(capi.lib().moog_env_post_reset(
                self.dev_program.handle, ctypes.byref(st), self.n, _ptr(rn), self._stream()))

    def env_step(self, actions=None, noise=None, rule_noise=None, auto_reset=True,
                 reset_index=None, want_counters=False, sample_resets=False, frames=None):
        """`moog_env_step`.  frames: None, True (draw the frame of the state every env is left in
        into `self.frames`) or a uint8 [N, H, W, 3] tensor to draw into -- on this device, or a
        pinned CPU tensor, which the kernels then write over PCIe as the envs finish."""
        p = self.program
        if frames is True:
            frames = self.frames
        if frames is not None:
            if self.frames is None:
                raise capi.MoogError('the program has no PILRenderer observer')
            if (frames.dtype != torch.uint8 or tuple(frames.shape) != tuple(self.frames.shape)
                    or not frames.is_contiguous()):
                raise ValueError('frames must be a contiguous uint8 tensor of shape {}'.format(
                    tuple(self.frames.shape)))
            if frames.device.type == 'cpu':
                if frames.data_ptr() not in self._pinned_ok:
                    if not frames.is_pinned():
                        raise ValueError('host frames must live in pinned memory')
                    self._pinned_ok.add(frames.data_ptr())
            elif frames.device != self.state.dyn.device:
                raise ValueError('frames are on {}, the envs on {}'.format(frames.device, self.state.dyn.device))
        act = self._as_f64(actions, max(p.action_dim, 1)) if actions is not None else None
        nz = self._as_f64(noise, p.K * p.noise_dim) if (noise is not None and p.noise_dim) else None
        rn = self._as_f64(rule_noise, p.rule_noise_dim) if (
            rule_noise is not None and p.rule_noise_dim) else None
        ri = None
        if reset_index is not None:
            ri = torch.as_tensor(reset_index).to(self.device, dtype=torch.int32).contiguous()
        io = capi.MoogStepIO()
        io.actions, io.noise, io.rule_noise = _ptr(act), _ptr(nz), _ptr(rn)
        pool_struct = None
        if auto_reset and self.pool is not None:
            pool_struct = self.pool.struct()
            io.pool = ctypes.pointer(pool_struct)
            io.pool_size = self.pool.n
        io.reset_index = _ptr(ri)
        io.sample_resets = 1 if sample_resets else 0
        io.seed = self.call_seed()
        io.reward, io.step_type, io.discount = _ptr(self.reward), _ptr(self.step_type), _ptr(self.discount)
        io.counters = _ptr(self.counters) if want_counters else None
        io.stats = _ptr(self.stats)
        io.frames = _ptr(frames)
        st = self.state.struct()
        with self._on_device():
            capi.check(capi.lib().moog_env_step(
                self.dev_program.handle, ctypes.byref(st), self.n, ctypes.byref(io), self._stream()))
        self._calls += 1

    def physics_step(self, noise=None, want_counters=True):
        p = self.program
        nz = self._as_f64(noise, p.K * p.noise_dim) if (noise is not None and p.noise_dim) else None
        st = self.state.struct()
        with torch.cuda.device(self.device):
            capi.check(capi.lib().moog_physics_step(
                self.dev_program.handle, ctypes.byref(st), self.n, _ptr(nz),
                _ptr(self.counters) if want_counters else None, self._stream()))

    def overlap_pairs(self, layer_a, layer_b):
        p = self.program
        la, lb = p.layer_index(layer_a), p.layer_index(layer_b)
        out = torch.zeros((self.n, p.layer_cap[la], p.layer_cap[lb]), dtype=torch.uint8, device=self.device)
        st = self.state.struct()
        with torch.cuda.device(self.device):
            capi.check(capi.lib().moog_overlap_pairs(
                self.dev_program.handle, ctypes.byref(st), self.n, la, lb, _ptr(out), self._stream()))
        return out

    def render(self, out=None, first=0, count=None):
        """Frames of the envs [first, first + count) into out[first:first + count]."""
        if self.frames is None:
            raise capi.MoogError('the program has no PILRenderer observer')
        out = self.frames if out is None else out
        count = self.n - first if count is None else count
        if first < 0 or count < 0 or first + count > self.n:
            raise ValueError('env range [{}, {}) is outside [0, {})'.format(first, first + count, self.n))
        st = self.state.struct(first)
        dst = ctypes.c_void_p(out.data_ptr() + first * out.stride(0) * out.element_size())
        with torch.cuda.device(self.device):
            capi.check(capi.lib().moog_render(
                self.dev_program.handle, ctypes.byref(st), count, dst, self._stream()))
        return out


class BatchedEnvironment(object):
    """See module docstring.  Reference: moog/environment.py:28-131."""

    def __init__(self, state_initializer, physics, task, action_space, observers,
                 game_rules=(), meta_state_initializer=None, *, num_envs,
                 device='cuda', pool_size=None, seed=0, layer_capacity=None,
                 initial_states=None, reset_mode='pool'):
        # (meta_state: None, or a dict of numbers / strings whose entries become variables of the env's
        # record -- compile_config refuses anything else)
        self.state_initializer = state_initializer
        self.num_envs = int(num_envs)
        config = dict(state_initializer=state_initializer, physics=physics, task=task,
                      action_space=action_space, observers=observers, game_rules=game_rules,
                      meta_state_initializer=meta_state_initializer)
        if initial_states is None:
            pool_size = int(pool_size or min(self.num_envs, 256))
            initial_states = [state_initializer() for _ in range(pool_size)]
        if reset_mode not in ('pool', 'device'):
            raise ValueError("reset_mode must be 'pool' or 'device'")
        self.reset_mode = reset_mode
        self.program = compiler.compile_config(config, initial_states, layer_capacity,
                                               reset_sampler=(reset_mode == 'device'))
        if reset_mode == 'device':
            # pool entry 0 is the template the device sampler starts every reset from
            initial_states = [self.program.reset_template] + list(initial_states)
        self._pool_arrays = compiler.pack_states(self.program, initial_states)
        self.engine = Engine(self.program, self.num_envs, device, seed)
        self.engine.set_pool(self._pool_arrays)
        self.action_space = action_space
        self.observers = observers
        self._image_key = None
        for k, obs in (observers or {}).items():
            if type(obs).__name__ == 'PILRenderer':
                self._image_key = k
        self._rng = np.random.RandomState(seed)
        self._started = False
        self._copy_stream = None
        self._copy_events = None
        self._auto_choice = None
        self._auto_calls = 0
        self._auto_times = {'mapped': [], 'chunked': []}

    # -- dm_env-like protocol -------------------------------------------------
    def _clear_envi(self):
        """Counters and error words back to 0 -- but not the two draw serials (MOOG_EI_CREATED,
        MOOG_EI_RULE_PASSES): they key the Philox draws of CreateSprites / random conditions, and an
        env reset by hand must not see the draws of its first episode again."""
        envi = self.engine.state.envi
        envi[:, :4].zero_()
        envi[:, 6:].zero_()

    def reset(self):
        """environment.py:82-96 for every env."""
        e = self.engine
        if self.reset_mode == 'device':
            # every env is marked terminated; the step kernel then resets it (environment.py:100-101)
            # and draws its generated sprites on the device
            for k in DeviceState.KEYS:
                getattr(e.state, k).copy_(getattr(e.pool, k)[0:1].expand_as(getattr(e.state, k)))
            self._clear_envi()
            e.state.envi[:, 1] = 1
            e.env_step(None, sample_resets=True)
            self._started = True
            return self._timestep()
        idx = torch.from_numpy(self._rng.randint(0, e.pool.n, size=self.num_envs)).to(e.device)
        for k in DeviceState.KEYS:
            if k == 'envi':
                self._clear_envi()
            else:
                getattr(e.state, k).copy_(getattr(e.pool, k).index_select(0, idx))
        e.post_reset()
        e.step_type.fill_(STEP_FIRST)
        e.reward.fill_(float('nan'))
        e.discount.fill_(float('nan'))
        self._started = True
        return self._timestep()

    def step(self, action=None):
        """environment.py:98-126 for every env; `action` is [N, action_dim]
        (Joystick: 2 floats, Grid: 1 index; Composite: dict of those or the
        concatenation in the order of `program.action_layout`)."""
        if not self._started:
            return self.reset()
        # the observation is drawn by the step call itself (inside the step kernel when a canvas
        # fits next to the env record, see include/moog_b200.h `frames`)
        self.engine.env_step(self._flatten_action(action), sample_resets=(self.reset_mode == 'device'),
                             frames=True if self._image_key is not None else None)
        return self._timestep(rendered=True)

    def step_to_host(self, action, host, chunks=4, frames='auto'):
        """`step(action)` whose TimeStep lands in the caller's pinned host buffers:
        `host` is a TimeStep of CPU tensors (step_type int32[N], reward float32[N],
        discount float32[N] or None, observation {'image': uint8[N,H,W,3]}).  Returns
        `host` once everything has arrived (the caller owns the TimeStep).

        frames = 'mapped': the step call draws every env's frame straight into the pinned host
        image (the kernels' stores cross PCIe while other envs are still being stepped; no
        device copy of the frames exists).  'device': the step call draws into device memory,
        one device-to-host copy follows.  'chunked': the step, then the frames rendered in
        `chunks` env ranges, the device-to-host copy of one range running on a second stream
        while the next range is rendered.  'auto': with a pinned host image the first 8 calls
        alternate between 'mapped' and 'chunked', time themselves, and the faster one is kept
        (mapped stores hide behind the step but move fewer bytes per second than the copy
        engines); 'device' when the image is not pinned."""
        img = host.observation[self._image_key] if self._image_key is not None else None
        auto = frames == 'auto'
        if auto:
            frames = self._auto_frames(img)
        t_begin = time.perf_counter() if auto else 0.0
        if frames not in ('mapped', 'device', 'chunked'):
            raise ValueError("frames must be 'auto', 'mapped', 'device' or 'chunked'")
        e = self.engine
        main = torch.cuda.current_stream(e.device)
        if frames != 'chunked' or not self._started:
            if not self._started:
                self.reset()
                if img is not None:
                    img.copy_(e.frames, non_blocking=True)
            else:
                dst = None if img is None else (img if frames == 'mapped' else True)
                e.env_step(self._flatten_action(action), sample_resets=(self.reset_mode == 'device'), frames=dst)
                if img is not None and frames == 'device':
                    img.copy_(e.frames, non_blocking=True)
            host.step_type.copy_(e.step_type, non_blocking=True)
            host.reward.copy_(e.reward, non_blocking=True)
            if host.discount is not None:
                host.discount.copy_(e.discount, non_blocking=True)
            main.synchronize()
            if auto:
                self._auto_record(frames, time.perf_counter() - t_begin)
            return host
        self.engine.env_step(self._flatten_action(action), sample_resets=(self.reset_mode == 'device'))
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=e.device)
            self._copy_events = [torch.cuda.Event() for _ in range(64)]
        copy = self._copy_stream
        host.step_type.copy_(e.step_type, non_blocking=True)
        host.reward.copy_(e.reward, non_blocking=True)
        if host.discount is not None:
            host.discount.copy_(e.discount, non_blocking=True)
        if self._image_key is not None:
            dst = host.observation[self._image_key]
            chunks = max(1, min(int(chunks), len(self._copy_events), self.num_envs))
            # equal ranges: measured on B200 / PCIe 5 with 4096 64x64 frames, 4 equal ranges beat
            # 1, 2, 8 or 16 ranges and growing / shrinking ones (each render launch has a fixed
            # cost and does not fill the GPU when it is small)
            first = 0
            per = (self.num_envs + chunks - 1) // chunks
            for c in range(chunks):
                count = per
                count = min(count, self.num_envs - first)
                if count <= 0:
                    continue
                e.render(first=first, count=count)
                ev = self._copy_events[c]
                ev.record(main)
                copy.wait_event(ev)
                with torch.cuda.stream(copy):
                    dst[first:first + count].copy_(e.frames[first:first + count], non_blocking=True)
                first += count
        copy.synchronize()
        main.synchronize()
        if auto:
            self._auto_record(frames, time.perf_counter() - t_begin)
        return host

    def host_pipeline(self, depth=2, frames='mapped'):
        """A `HostPipeline` over this environment: `submit(action)` / `collect()` with `depth`
        steps in flight, every TimeStep delivered in pinned host memory."""
        return HostPipeline(self, depth=depth, frames=frames)

    # frames='auto': kernel stores into the mapped host image hide the transfer behind the step
    # when the frames are small next to the step (4096 64x64 frames: 50 MB per 4.5 ms), but they
    # move fewer bytes per second over PCIe than the copy engines (16384 84x84 frames, 347 MB per
    # step: 986 k env-steps/s mapped, 1.48 M chunked).  The first calls try both and time them.
    _AUTO_TRIALS = 8

    def _auto_frames(self, img):
        if img is None or not img.is_pinned():
            return 'device'
        if self._auto_choice is not None:
            return self._auto_choice
        return ('mapped', 'chunked')[self._auto_calls % 2]

    def _auto_record(self, mode, seconds):
        if self._auto_choice is not None or mode not in self._auto_times:
            return
        self._auto_calls += 1
        if self._auto_calls > 2:                 # the first call of each mode warms up
            self._auto_times[mode].append(seconds)
        if self._auto_calls >= self._AUTO_TRIALS:
            best = {m: min(v) for m, v in self._auto_times.items() if v}
            self._auto_choice = min(best, key=best.get) if best else 'mapped'

    def observation(self):
        obs = {}
        if self._image_key is not None:
            obs[self._image_key] = self.engine.render()
        return obs

    def _timestep(self, rendered=False):
        e = self.engine
        if rendered:
            obs = {self._image_key: e.frames} if self._image_key is not None else {}
        else:
            obs = self.observation()
        return TimeStep(e.step_type, e.reward, e.discount, obs)

    def _flatten_action(self, action):
        if action is None:
            return None
        if isinstance(action, dict):
            cols = []
            for key, _, _, width in self.program.action_layout:
                a = action[key]
                a = a if torch.is_tensor(a) else torch.as_tensor(np.asarray(a))
                cols.append(a.to(self.engine.device, dtype=torch.float64).reshape(self.num_envs, width))
            return torch.cat(cols, dim=1)
        return action

    # -- extras -----------------------------------------------------------------
    def meta_state(self, env=0):
        """The env's meta_state as the dict the reference would hold (environment.py:129-131): the entries the
        program carries, strings decoded; None when the config has no meta_state."""
        prog = self.program
        if not getattr(prog, 'meta_vars', None):
            return None
        row = self.engine.state.envf[env].cpu().numpy()
        out = {}
        for key, slot in prog.meta_vars.items():
            v = float(row[slot])
            if isinstance(prog.meta_var_init.get(key), str) or key in getattr(prog, 'string_vars', ()):
                out[key] = prog.strings[int(v) - 1] if 1 <= int(v) <= len(prog.strings) else ''
            else:
                out[key] = int(v) if v == int(v) and abs(v) < 2 ** 53 else v
        return out

    def raise_errors(self):
        """Raises the data-dependent exceptions the reference would have raised."""
        err = self.engine.state.envi[:, 2]  # MOOG_EI_ERR
        bad = torch.nonzero(err).flatten()
        if bad.numel():
            i = int(bad[0])
            code = int(err[i])
            text = '; '.join(t for b, t in _ERR_TEXT.items() if code & b)
            raise ValueError('env {}: {}'.format(i, text))

    def state_dict(self):
        """Snapshot (cf. SimulationEnvironment, env_wrappers/simulation.py:55-83)."""
        return {k: getattr(self.engine.state, k).clone() for k in DeviceState.KEYS}

    def load_state_dict(self, sd):
        for k in DeviceState.KEYS:
            getattr(self.engine.state, k).copy_(sd[k])

    def episode_stats(self, reduce=True):
        """[sum reward, sum finished-episode length, finished episodes, env-steps];
        summed over ranks with one NCCL all_reduce when torch.distributed is up."""
        from . import dist as mdist
        s = self.engine.stats.clone()
        return mdist.reduce_stats(s) if reduce else s

    @property
    def action_dim(self):
        return max(self.program.action_dim, 1)


class HostPipeline(object):
    """Host-facing stepping with `depth` steps in flight (double buffering at depth 2).

    `step_to_host` hands the host one finished TimeStep per call and the GPU idles while the host
    looks at it (synchronise, Python, the next call's enqueue) -- and 8 ranks doing that against
    one host lose 20 % (round-1 scaling run).  Here the caller submits the action of step k+1
    BEFORE it collects the TimeStep of step k:

        pipe = env.host_pipeline(depth=2)
        pipe.submit(a0)
        for k in range(1, T):
            pipe.submit(a[k])          # enqueued behind step k-1, returns at once
            ts = pipe.collect()        # TimeStep of step k-1, in this slot's pinned host buffers
        ts = pipe.collect()

    i.e. the policy acts on an observation that is one step old (the usual actor pipeline); with
    depth 1 it is `step_to_host`.  Every submit copies that step's actions host -> device and
    every collect returns step_type / reward / discount / frames in pinned host memory, valid
    until `depth` further submits.  frames = 'mapped': the kernels store the frames straight
    into the slot's pinned image over PCIe as the envs finish; 'device': they are drawn in HBM
    and copied by a copy engine on a second stream while the next step runs.
    Reference protocol: moog/environment.py:98-126 (one `step` per submit, auto-reset included).
    """

    def __init__(self, env, depth=2, frames='mapped'):
        if depth < 1:
            raise ValueError('depth must be >= 1')
        if frames not in ('mapped', 'device'):
            raise ValueError("frames must be 'mapped' or 'device'")
        self.env, self.depth, self.frames = env, int(depth), frames
        e = env.engine
        self._main = torch.cuda.current_stream(e.device)
        self._copy = torch.cuda.Stream(device=e.device) if frames == 'device' else None
        n, ad = env.num_envs, env.action_dim
        pin = lambda *shape, dtype: torch.empty(shape, dtype=dtype).pin_memory()
        self._slots = []
        for _ in range(self.depth):
            img = None
            if env._image_key is not None:  # pylint: disable=protected-access
                img = pin(*e.frames.shape, dtype=torch.uint8)
            self._slots.append(dict(
                ts=TimeStep(pin(n, dtype=torch.int32), pin(n, dtype=torch.float32), pin(n, dtype=torch.float32),
                            {env._image_key: img} if img is not None else {}),  # pylint: disable=protected-access
                img=img,
                dev_frames=(torch.empty_like(e.frames) if (img is not None and frames == 'device') else None),
                host_act=pin(n, ad, dtype=torch.float64),
                dev_act=torch.empty((n, ad), dtype=torch.float64, device=e.device),
                done=torch.cuda.Event(), stepped=torch.cuda.Event(), busy=False))
        self._head = 0     # next slot to submit into
        self._tail = 0     # next slot to collect
        self._in_flight = 0

    def submit(self, action=None):
        """Enqueue one `Environment.step(action)`; returns immediately."""
        if self._in_flight >= self.depth:
            raise RuntimeError('pipeline full: collect() a TimeStep before submitting another step')
        env, e = self.env, self.env.engine
        slot = self._slots[self._head]
        was_reset = not env._started  # pylint: disable=protected-access
        with torch.cuda.stream(self._main):
            if was_reset:
                env.reset()                                   # the first step of a MOOG env is its reset
                if slot['img'] is not None:
                    slot['img'].copy_(e.frames, non_blocking=True)
            else:
                act = None
                if action is not None:
                    action = env._flatten_action(action)  # pylint: disable=protected-access
                    if torch.is_tensor(action) and action.is_cuda:
                        slot['dev_act'].copy_(action.reshape(slot['dev_act'].shape), non_blocking=True)
                    else:
                        # staged in this slot's pinned buffer: the caller may reuse its own at once
                        slot['host_act'].copy_(torch.as_tensor(action).reshape(slot['host_act'].shape))
                        slot['dev_act'].copy_(slot['host_act'], non_blocking=True)
                    act = slot['dev_act']
                dst = None
                if slot['img'] is not None:
                    dst = slot['img'] if self.frames == 'mapped' else slot['dev_frames']
                e.env_step(act, sample_resets=(env.reset_mode == 'device'), frames=dst)
            ts = slot['ts']
            ts.step_type.copy_(e.step_type, non_blocking=True)
            ts.reward.copy_(e.reward, non_blocking=True)
            ts.discount.copy_(e.discount, non_blocking=True)
            if self.frames == 'device' and slot['img'] is not None and not was_reset:
                slot['stepped'].record(self._main)
                self._copy.wait_event(slot['stepped'])
                with torch.cuda.stream(self._copy):
                    slot['img'].copy_(slot['dev_frames'], non_blocking=True)
                    slot['done'].record(self._copy)
            else:
                slot['done'].record(self._main)
        slot['busy'] = True
        self._head = (self._head + 1) % self.depth
        self._in_flight += 1

    def collect(self):
        """The oldest submitted step's TimeStep (host tensors), once all of it has arrived."""
        if self._in_flight == 0:
            raise RuntimeError('nothing in flight')
        slot = self._slots[self._tail]
        slot['done'].synchronize()
        slot['busy'] = False
        self._tail = (self._tail + 1) % self.depth
        self._in_flight -= 1
        return slot['ts']

    def last_event(self):
        """CUDA event of the most recently submitted step's completion (for device-side timing)."""
        return self._slots[(self._head - 1) % self.depth]['done']

    def step(self, action=None):
        """submit + collect at depth 1 semantics: the TimeStep of THIS action (drains the pipeline first)."""
        while self._in_flight:
            self.collect()
        self.submit(action)
        return self.collect()
