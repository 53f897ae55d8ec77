"""`Physics.step(state)` called on the HOST by a config (abstract_physics.py:39-42) -- e.g. the
state initializers of red_green.py:87-101 and bounce_box_contact_prediction.py:26-36 roll their
physics forward to predict how a trial will end before they accept an initial state.

There is no CPU physics in this package: the call packs the host sprites into a batch of ONE env,
runs the same CUDA physics kernel the batched environment runs (`moog_physics_step`: K substeps of
forces, collisions and position updates, physics.py:60-86), and writes positions, velocities, angles
and angular velocities back into the host `Sprite` objects.  It therefore needs the GPU, like
everything else here, and costs a launch per call (a pool of initial states is built once).

RandomForce / RandomMazeWalk entries draw their uniforms from `np.random`, like the reference.
"""

import numpy as np

from . import compiler

_PROGRAMS = {}      # id(physics) -> (layout key, physics, Program, Engine)


def _layout(state):
    return tuple((name, tuple(len(sp.vertices) for sp in sprites)) for name, sprites in state.items())


def _program_for(physics, state):
    import torch
    from moog import action_spaces, tasks
    from .batched_env import Engine
    from . import capi
    if not torch.cuda.is_available():
        raise capi.MoogError('Physics.step on the host runs the CUDA physics kernel: it needs a CUDA device '
                             '(there is no CPU path)')
    key = _layout(state)
    cached = _PROGRAMS.get(id(physics))
    if cached is not None and cached[0] == key and cached[1] is physics:
        return cached[2], cached[3]
    config = dict(physics=physics, task=tasks.CompositeTask(), action_space=action_spaces.Grid(action_layers=()),
                  observers={}, game_rules=())
    prog = compiler.compile_config(config, [state])
    engine = Engine(prog, 1, 'cuda')
    _PROGRAMS[id(physics)] = (key, physics, prog, engine)
    return prog, engine


def step(physics, state):
    """One `physics.step(state)`: updates_per_env_step substeps, in place on the host sprites."""
    prog, engine = _program_for(physics, state)
    arrays = compiler.pack_states(prog, [state])
    engine.state.upload({k: arrays[k] for k in ('dyn', 'stat', 'meta', 'vtx', 'cnt', 'envi', 'envf')})
    noise = None
    if prog.noise_dim:
        noise = np.random.uniform(size=(1, prog.K * prog.noise_dim))
    engine.physics_step(noise=noise, want_counters=False)
    out = engine.state.download()
    err = int(out['envi'][0, 2])
    if err:
        from .batched_env import _ERR_TEXT
        raise ValueError('; '.join(t for b, t in _ERR_TEXT.items() if err & b))
    for l, name in enumerate(prog.layer_names):
        for k, sp in enumerate(state[name]):
            s = prog.layer_off[l] + k
            d = out['dyn'][0][:, s]
            flags = int(out['meta'][0][1, s])
            vel = np.array([d[2], d[3]])
            if flags & compiler.SF_VEL32:
                vel = vel.astype(np.float32)
            sp.position = np.array([d[0], d[1]])
            sp.velocity = vel
            sp.angle = _as_kind(d[4], (flags >> compiler.SF_ANG_SHIFT) & 3)
            sp.angle_vel = _as_kind(d[5], (flags >> compiler.SF_ANGVEL_SHIFT) & 3)


def _as_kind(value, kind):
    """The scalar with the NumPy kind the record tracks for it (0 python float, 1 float32, 2 float64)."""
    if kind == 1:
        return np.float32(value)
    if kind == 2:
        return np.float64(value)
    return float(value)
