"""`dm_env.specs` stand-in: Array / BoundedArray / DiscreteArray.

TEST INFRASTRUCTURE ONLY (see dm_env/__init__.py).  Call sites in the
reference: observers/pil_renderer.py:78, action_spaces/joystick.py:42,
action_spaces/grid.py:50, action_spaces/set_position.py:31.
"""

import numpy as np


class Array:
    def __init__(self, shape, dtype, name=None):
        self.shape = tuple(int(d) for d in shape)
        self.dtype = np.dtype(dtype)
        self.name = name

    def __repr__(self):
        return 'Array(shape={}, dtype={}, name={})'.format(
            self.shape, self.dtype, self.name)

    def generate_value(self):
        return np.zeros(self.shape, self.dtype)


class BoundedArray(Array):
    def __init__(self, shape, dtype, minimum, maximum, name=None):
        super().__init__(shape, dtype, name)
        self.minimum = np.array(minimum, dtype=self.dtype)
        self.maximum = np.array(maximum, dtype=self.dtype)

    def generate_value(self):
        return (np.ones(self.shape, self.dtype) * self.minimum).astype(
            self.dtype)


class DiscreteArray(BoundedArray):
    def __init__(self, num_values, dtype=np.int32, name=None):
        super().__init__((), dtype, 0, num_values - 1, name)
        self.num_values = int(num_values)
