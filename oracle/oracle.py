"""ctypes front-end of the CPU oracle (oracle/moog_oracle.c, pil_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product package.

The oracle consumes the same program blob and state-record arrays as the CUDA
library (include/moog_b200_program.h), one env at a time, in float64.
"""

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libmoog_oracle.so')

MAX_LAYERS = 16
MAX_VERTS = 128   # MOOG_MAX_OUTLINE


def build(force=False):
    """Compile the oracle with gcc (make -C oracle)."""
    srcs = [os.path.join(_HERE, f) for f in ('moog_oracle.c', 'pil_oracle.c')]
    srcs.append(os.path.join(_HERE, '..', 'include', 'moog_b200_program.h'))
    stale = force or not os.path.exists(_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return _SO


class _State(ctypes.Structure):
    _fields_ = [('dyn', ctypes.c_void_p), ('stat', ctypes.c_void_p),
                ('meta', ctypes.c_void_p), ('cnt', ctypes.c_void_p),
                ('envi', ctypes.c_void_p), ('envf', ctypes.c_void_p),
                ('vtx', ctypes.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


class Oracle(object):
    """Steps a batch of env records on the CPU, in place."""

    def __init__(self, program, arrays):
        """program: moog_b200.compiler.Program; arrays: pack_states() dict
        (the arrays are copied; the oracle owns its state)."""
        self.program = program
        self.blob = np.frombuffer(program.blob, dtype=np.uint8).copy()
        self.dyn = _c(arrays['dyn'], np.float64).copy()
        self.stat = _c(arrays['stat'], np.float64).copy()
        self.meta = _c(arrays['meta'], np.int32).copy()
        self.cnt = _c(arrays['cnt'], np.int32).copy()
        self.envi = _c(arrays['envi'], np.int32).copy()
        self.envf = _c(arrays['envf'], np.float64).copy()
        self.vtx = _c(arrays['vtx'], np.float64).copy()
        self.n = self.dyn.shape[0]
        self.counters = np.zeros((self.n, 4), dtype=np.int64)

    def arrays(self):
        return dict(dyn=self.dyn, stat=self.stat, meta=self.meta, vtx=self.vtx,
                    cnt=self.cnt, envi=self.envi, envf=self.envf)

    def _state(self):
        return _State(_ptr(self.dyn), _ptr(self.stat), _ptr(self.meta),
                      _ptr(self.cnt), _ptr(self.envi), _ptr(self.envf),
                      _ptr(self.vtx))

    def post_reset(self, rule_noise=None):
        rn = _c(rule_noise, np.float64)
        st = self._state()
        lib().orc_env_post_reset(_ptr(self.blob), ctypes.byref(st),
                                 ctypes.c_int(self.n), _ptr(rn))

    def physics_step(self, noise=None):
        nz = _c(noise, np.float64)
        st = self._state()
        lib().orc_physics_step(_ptr(self.blob), ctypes.byref(st),
                               ctypes.c_int(self.n), _ptr(nz),
                               _ptr(self.counters))

    def step(self, actions, noise=None, rule_noise=None):
        """environment.py:98-126 for every env. Returns (reward, step_type)."""
        ad = max(self.program.action_dim, 1)
        act = np.zeros((self.n, ad)) if actions is None else _c(
            np.asarray(actions, dtype=np.float64).reshape(self.n, ad),
            np.float64)
        nz = _c(noise, np.float64)
        rn = _c(rule_noise, np.float64)
        reward = np.zeros(self.n)
        step_type = np.zeros(self.n, dtype=np.int32)
        st = self._state()
        lib().orc_env_step(
            _ptr(self.blob), ctypes.byref(st), ctypes.c_int(self.n), _ptr(act),
            _ptr(nz), _ptr(rn), ctypes.c_int(self.program.rule_noise_dim),
            _ptr(reward), _ptr(step_type), _ptr(self.counters))
        return reward, step_type

    @staticmethod
    def set_seed(seed):
        """Key of the counter-based draws (CreateSprites, random conditions without a rule-noise
        tensor) of the calls that follow: what the CUDA path receives as io.seed."""
        lib().orc_set_seed(ctypes.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF))

    @staticmethod
    def set_sample_resets(on):
        """step_auto draws the generate_sprites groups of the program's reset section per env
        (moog_step_io.sample_resets) instead of copying a pool row."""
        lib().orc_set_sample_resets(ctypes.c_int(1 if on else 0))

    _forced = None

    @classmethod
    def force_factors(cls, rows):
        """The next generate_sprites tries take these factor rows ([n, 14]) instead of drawing
        (replay of a recorded reference trajectory); None / empty turns the hook off."""
        rows = None if rows is None or len(rows) == 0 else np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, 14)
        cls._forced = rows          # keeps the buffer alive
        lib().orc_force_factors(_ptr(rows), ctypes.c_int(0 if rows is None else len(rows)))

    @staticmethod
    def forced_left():
        return int(lib().orc_forced_left())

    def step_auto(self, actions, pool, reset_index, noise=None, rule_noise=None):
        """Environment.step with its auto-reset (environment.py:98-126 incl. :100-101,
        reset() :82-96): envs whose previous transition terminated take row
        reset_index[n] of `pool` (another Oracle holding initial states) as the
        state initializer's result.  Returns (reward, step_type, discount)."""
        ad = max(self.program.action_dim, 1)
        act = np.zeros((self.n, ad)) if actions is None else _c(
            np.asarray(actions, dtype=np.float64).reshape(self.n, ad),
            np.float64)
        nz = _c(noise, np.float64)
        rn = _c(rule_noise, np.float64)
        ri = _c(np.asarray(reset_index).reshape(self.n), np.int32)
        reward = np.zeros(self.n)
        discount = np.zeros(self.n)
        step_type = np.zeros(self.n, dtype=np.int32)
        st, pst = self._state(), pool._state()
        lib().orc_env_step_auto(
            _ptr(self.blob), ctypes.byref(st), ctypes.c_int(self.n),
            ctypes.byref(pst), ctypes.c_int(pool.n), _ptr(ri), _ptr(act),
            _ptr(nz), _ptr(rn), ctypes.c_int(self.program.rule_noise_dim),
            _ptr(reward), _ptr(step_type), _ptr(discount), _ptr(self.counters))
        return reward, step_type, discount

    def overlap_pairs(self, layer_a, layer_b):
        p = self.program
        la, lb = p.layer_index(layer_a), p.layer_index(layer_b)
        out = np.zeros((self.n, p.layer_cap[la], p.layer_cap[lb]),
                       dtype=np.uint8)
        st = self._state()
        lib().orc_overlap_pairs(_ptr(self.blob), ctypes.byref(st),
                                ctypes.c_int(self.n), ctypes.c_int(la),
                                ctypes.c_int(lb), _ptr(out))
        return out

    def world_vertices(self):
        S = self.program.n_slots
        out = np.zeros((self.n, S, MAX_VERTS, 2))
        nv = np.zeros((self.n, S), dtype=np.int32)
        st = self._state()
        lib().orc_world_vertices(_ptr(self.blob), ctypes.byref(st),
                                 ctypes.c_int(self.n), _ptr(out), _ptr(nv))
        return out, nv

    def render(self):
        """PILRenderer.__call__ restated in C -> uint8 [n, H, W, 3]."""
        r = self.program.render
        out = np.zeros((self.n, r['height'], r['width'], 3), dtype=np.uint8)
        st = self._state()
        lib().orc_render(_ptr(self.blob), ctypes.byref(st),
                         ctypes.c_int(self.n), _ptr(out))
        return out
