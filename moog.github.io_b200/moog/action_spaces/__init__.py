"""Action-space specs (reference: moog/action_spaces/{joystick,grid,
set_position,composite}.py).  `random_action` runs on the host for
single-env convenience; `step` happens inside the device physics kernel."""

import abc

import numpy as np


def _as_tuple(x):
    return tuple(x) if isinstance(x, (list, tuple)) else (x,)


class _Spec(object):
    """Tiny array-spec record (stands in for dm_env.specs)."""

    def __init__(self, shape, dtype, minimum=None, maximum=None,
                 num_values=None):
        self.shape = shape
        self.dtype = np.dtype(dtype)
        self.minimum = minimum
        self.maximum = maximum
        self.num_values = num_values


class AbstractActionSpace(abc.ABC):
    def reset(self, state):
        pass

    def step(self, state, action):
        raise RuntimeError(
            'action spaces are applied on the device by BatchedEnvironment')

    def action_spec(self):
        return self._action_spec


class Joystick(AbstractActionSpace):
    """2-D continuous force / velocity control (joystick.py:12-43)."""

    def __init__(self, scaling_factor=1., action_layers='agent',
                 constrained_lr=False, control_velocity=False, momentum=0.):
        self._scaling_factor = scaling_factor
        self._action_layers = _as_tuple(action_layers)
        self._constrained_lr = constrained_lr
        self._control_velocity = control_velocity
        self._momentum = momentum
        self._action_spec = _Spec((2,), np.float32, -1, 1)

    def random_action(self):
        return np.random.uniform(-1., 1., size=(2,))


class Grid(AbstractActionSpace):
    """5 discrete actions: left, right, down, up, stay (grid.py:11-50)."""

    _ACTIONS = (
        np.array([-1, 0]), np.array([1, 0]), np.array([0, -1]),
        np.array([0, 1]), np.array([0, 0]),
    )

    def __init__(self, scaling_factor=1., action_layers='agent',
                 control_velocity=False, momentum=0.):
        self._scaling_factor = scaling_factor
        self._action_layers = _as_tuple(action_layers)
        self._control_velocity = control_velocity
        self._momentum = momentum
        self._action_spec = _Spec((), np.int32, 0, 4, num_values=5)

    def random_action(self):
        return np.random.randint(len(Grid._ACTIONS))


class SetPosition(AbstractActionSpace):
    """Sets sprite positions directly, with inertia (set_position.py:14-32)."""

    def __init__(self, action_layers='agent', inertia=0.):
        self._action_layers = _as_tuple(action_layers)
        self._inertia = inertia
        self._action_spec = _Spec((2,), np.float32, 0, 1)

    def random_action(self):
        return np.random.uniform(0., 1., size=(2,))


class Composite(AbstractActionSpace):
    """Dict of named action spaces (composite.py:11-69)."""

    def __init__(self, **action_spaces):
        self.action_spaces = action_spaces
        self._action_keys = action_spaces.keys()
        self._action_spec = {
            k: v.action_spec() for k, v in action_spaces.items()}

    def random_action(self):
        return {k: self.action_spaces[k].random_action()
                for k in self._action_keys}

    @property
    def action_keys(self):
        return list(self._action_keys)
