"""Batched counterparts of the reference's env wrappers (moog/env_wrappers/).

`VectorGymWrapper` gives a `BatchedEnvironment` the (obs, reward, done, info)
protocol of the reference's `GymWrapper` (gym_wrapper.py:48-140), one entry per
env, without depending on `gym`: spaces are described by plain dicts.
`BatchedSimulation` is the snapshot / restore facility of
`SimulationEnvironment` (simulation.py:55-83) for tree search: the whole batch
is saved to and restored from device tensors.

Host-side convenience only; the step itself is `BatchedEnvironment.step`.
"""
import torch

from .batched_env import STEP_LAST


class VectorGymWrapper(object):
    """Gym-style vector env over a BatchedEnvironment (auto-resetting)."""

    metadata = {'render.modes': ['rgb_array']}

    def __init__(self, env):
        self._env = env
        self._last_render = None
        self._env.reset()

    @property
    def num_envs(self):
        return self._env.num_envs

    @property
    def observation_space(self):
        """{key: dict(low, high, shape, dtype)} per env (gym_wrapper.py:66-78)."""
        spaces = {}
        r = self._env.program.render
        if r is not None and self._env._image_key is not None:  # pylint: disable=protected-access
            spaces[self._env._image_key] = dict(  # pylint: disable=protected-access
                low=0, high=255, shape=(r['height'], r['width'], 3), dtype='uint8')
        return spaces

    @property
    def action_space(self):
        """One entry per action-space component, in `program.action_layout` order:
        Joystick / SetPosition -> Box(-1, 1, (2,)) resp. Box(0, 1, (2,)), Grid -> Discrete(5)."""
        out = []
        for key, kind, _, width in self._env.program.action_layout:
            if kind == 'Grid':
                out.append(dict(key=key, type='Discrete', n=5))
            else:
                low = 0. if kind == 'SetPosition' else -1.
                out.append(dict(key=key, type='Box', low=low, high=1., shape=(width,), dtype='float32'))
        return out

    def _obs(self, ts):
        obs = dict(ts.observation)
        if 'image' in obs:
            self._last_render = obs['image']
        return obs

    def reset(self):
        return self._obs(self._env.reset())

    def step(self, action):
        """-> obs {key: tensor[N, ...]}, reward float32[N] (0 where dm_env has None),
        done bool[N], info {'discount': float32[N]}.  An env that is done is reset by
        the next call (environment.py:100-101), which ignores that env's action."""
        ts = self._env.step(action)
        reward = torch.nan_to_num(ts.reward, nan=0.0)
        done = ts.step_type == STEP_LAST
        return self._obs(ts), reward, done, {'discount': ts.discount}

    def render(self, mode='rgb_array'):
        del mode
        return self._last_render

    def close(self):
        pass


class BatchedSimulation(object):
    """Snapshot / restore of a whole batch (simulation.py:55-83): `SimulationEnvironment`'s
    protocol -- `reset`, `step`, `sim_step`, `sim_pop(index)` -- plus explicit `push` / `pop`.
    A snapshot is every array of the state record (task / action-space / rule memory lives in
    it: `envf`, `envi`) and the last TimeStep's scalars."""

    def __init__(self, env):
        self._env = env
        self._stack = []

    def _snapshot(self):
        e = self._env.engine
        return (self._env.state_dict(), e.reward.clone(), e.step_type.clone(), e.discount.clone())

    def _restore(self, snap):
        sd, reward, step_type, discount = snap
        self._env.load_state_dict(sd)
        e = self._env.engine
        e.reward.copy_(reward)
        e.step_type.copy_(step_type)
        e.discount.copy_(discount)

    def push(self):
        self._stack.append(self._snapshot())
        return len(self._stack)

    def pop(self):
        if not self._stack:
            raise IndexError('no snapshot to restore')
        self._restore(self._stack.pop())

    def reset(self):
        """simulation.py:60-62"""
        self._stack = []
        return self._env.reset()

    def step(self, action):
        """simulation.py:64-69: a real step first discards every simulated one."""
        if self._stack:
            self.sim_pop(index=0)
        self._stack = []
        return self._env.step(action)

    def sim_step(self, action):
        """simulation.py:71-87: snapshot, then step.  The reference refuses to simulate across an
        episode boundary (returns None when the env is about to reset); for a batch that is
        `None` as soon as ANY env is pending a reset."""
        if bool((self._env.engine.state.envi[:, 1] != 0).any()):    # MOOG_EI_RESET_NEXT
            return None
        self._stack.append(self._snapshot())
        return self._env.step(action)

    def sim_pop(self, index=-1):
        """simulation.py:89-100: restore snapshot `index`, drop it and everything after it."""
        self._restore(self._stack[index])
        self._stack = self._stack[:index]

    @property
    def stack_depth(self):
        return len(self._stack)
