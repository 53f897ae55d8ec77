"""Task (reward / termination) specs.

Reference: moog/tasks/{abstract_task,contact_reward,reset,stay_alive,
composite_task}.py.  Parameter holders under the reference's attribute
names; evaluated on the device by the rules/reward kernel.
"""

import abc

import numpy as np


def _as_list(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


class AbstractTask(abc.ABC):
    def reset(self, state, meta_state):
        pass

    def reward(self, state, meta_state, step_count):
        raise RuntimeError(
            'tasks are evaluated on the device by BatchedEnvironment')


class ContactReward(AbstractTask):
    """Reward when a layers_0 sprite overlaps a layers_1 sprite
    (contact_reward.py:20-68)."""

    def __init__(self, reward_fn, layers_0, layers_1, condition=None,
                 reset_steps_after_contact=np.inf):
        self._reward_fn = reward_fn
        self._layers_0 = _as_list(layers_0)
        self._layers_1 = _as_list(layers_1)
        self._condition = condition
        self._reset_steps_after_contact = reset_steps_after_contact


class Reset(AbstractTask):
    """Reset `steps_after_condition` steps after `condition(state)` first
    holds (reset.py:20-46)."""

    def __init__(self, condition, reward_fn=None, steps_after_condition=np.inf):
        self._condition = condition
        self._reward_fn = reward_fn
        self._steps_after_condition = steps_after_condition


class StayAlive(AbstractTask):
    """Periodic reward (stay_alive.py:9-20)."""

    def __init__(self, reward_period, reward_value=1.):
        self._reward_period = reward_period
        self._reward_value = reward_value


class CompositeTask(AbstractTask):
    """Sum of sub-task rewards, OR of resets, plus a timeout
    (composite_task.py:17-30)."""

    def __init__(self, *tasks, timeout_steps=np.inf):
        self._tasks = tasks
        self._timeout_steps = timeout_steps
