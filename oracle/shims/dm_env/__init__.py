"""Minimal stand-in for the `dm_env` package (pinned dm_env==1.6 in the
reference's setup.py:39-46; not installed in this image, no wheel available).

TEST INFRASTRUCTURE ONLY.  It exists so that the *unmodified* reference under
/root/reference can be imported in the build container to generate golden
vectors (oracle/gen_golden.py) and to run the reference's own KATs.  Only the
surface the reference touches is provided (call sites:
moog/environment.py:8,11,96,124,126; moog/env_wrappers/logger.py:206).
"""

import abc
import enum
from typing import Any, NamedTuple

from . import specs  # noqa: F401


class StepType(enum.IntEnum):
    FIRST = 0
    MID = 1
    LAST = 2

    def first(self):
        return self is StepType.FIRST

    def mid(self):
        return self is StepType.MID

    def last(self):
        return self is StepType.LAST


class TimeStep(NamedTuple):
    step_type: Any
    reward: Any
    discount: Any
    observation: Any

    def first(self):
        return self.step_type == StepType.FIRST

    def mid(self):
        return self.step_type == StepType.MID

    def last(self):
        return self.step_type == StepType.LAST


class Environment(abc.ABC):
    """Abstract episode-loop interface (reset / step / specs)."""

    @abc.abstractmethod
    def reset(self):
        pass

    @abc.abstractmethod
    def step(self, action):
        pass

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def restart(observation):
    return TimeStep(StepType.FIRST, None, None, observation)


def transition(reward, observation, discount=1.0):
    return TimeStep(StepType.MID, reward, discount, observation)


def termination(reward, observation):
    return TimeStep(StepType.LAST, reward, 0.0, observation)


def truncation(reward, observation, discount=1.0):
    return TimeStep(StepType.LAST, reward, discount, observation)
