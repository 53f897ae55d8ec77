"""Physics specs: forces, the rotational Collision, corrective physics.

Reference: moog/physics/ (physics.py, abstract_force.py, collisions.py,
gravity.py, friction.py, random_force.py, distance_fn_force.py,
tether_physics.py, constant_speed.py).  These classes only record their
parameters under the reference's attribute names; the config compiler
(moog_b200.compiler) lowers them to device ops and the sm_100a physics kernel
executes them.  `Physics.step(state)` on the host (used by a few state initializers) runs that same
kernel on a batch of one env (moog_b200/host_physics.py); single forces cannot be stepped on the host.
"""

import abc

import numpy as np


def _as_list(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


class _DeviceOnly(object):
    def _host_call(self, *args, **kwargs):
        raise RuntimeError(
            '{} is a device-side spec: it is executed by the CUDA physics '
            'kernel of moog_b200.BatchedEnvironment, not on the host.'.format(
                type(self).__name__))


class AbstractForce(_DeviceOnly, abc.ABC):
    """Base of all forces (abstract_force.py:10-41)."""

    def step(self, *sprites, updates_per_env_step):
        self._host_call()

    def reset(self, state):
        pass


class AbstractNewtonianForce(AbstractForce):
    """Forces obeying dv = F / (m K) (abstract_force.py:44-74)."""


class AbstractPhysics(_DeviceOnly, abc.ABC):
    """abstract_physics.py:6-47."""

    def __init__(self, updates_per_env_step):
        self._updates_per_env_step = updates_per_env_step

    def apply_physics(self, state, updates_per_env_step):
        self._host_call()

    def reset(self, state):
        pass

    def step(self, state):
        """abstract_physics.py:39-42 called by a config on the host (a state initializer that rolls
        the physics forward): a batch of one env through the CUDA physics kernel, results written
        back into the host sprites (moog_b200/host_physics.py).  Needs the GPU."""
        from moog_b200 import host_physics
        host_physics.step(self, state)

    @property
    def updates_per_env_step(self):
        return self._updates_per_env_step


class Physics(AbstractPhysics):
    """Ordered force list + corrective physics + substep count
    (physics.py:15-86)."""

    def __init__(self, *forces, updates_per_env_step=1, corrective_physics=()):
        super().__init__(updates_per_env_step=updates_per_env_step)
        self._forces = forces
        if not isinstance(corrective_physics, (list, tuple)):
            corrective_physics = [corrective_physics]
        self._corrective_physics = corrective_physics


class Collision(AbstractForce):
    """collisions.py:466-492."""

    def __init__(self, elasticity=1., symmetric=False, update_angle_vel=True,
                 max_recursion_depth=0):
        self._elasticity = elasticity
        self._symmetric = symmetric
        self._update_angle_vel = update_angle_vel
        self._max_recursion_depth = max_recursion_depth


class DownGravity(AbstractNewtonianForce):
    """F = g m (0, 1) (gravity.py:13-23)."""

    def __init__(self, g=-1.):
        self._g = g


class Gravity(AbstractNewtonianForce):
    """Pairwise F = g m0 m1 dist * unit(diff) (gravity.py:32-60)."""

    def __init__(self, g=-1., symmetric=True):
        self._g = g
        self._symmetric = symmetric


class KineticFriction(AbstractNewtonianForce):
    """F = -c m unit(v) (friction.py:15-33)."""

    def __init__(self, coeff_friction=1.):
        self._coeff_friction = coeff_friction


class Drag(AbstractNewtonianForce):
    """F = -c m v (friction.py:36-56)."""

    def __init__(self, coeff_friction=1.):
        self._coeff_friction = coeff_friction


class RandomForce(AbstractNewtonianForce):
    """Uniform magnitude in [0, max], uniform direction (random_force.py)."""

    def __init__(self, max_force_magnitude):
        self._max_force_magnitude = max_force_magnitude


class _LinearForceFn(object):
    """distance -> zero_intercept + slope * distance, optionally zeroed on
    either side of the event horizon (distance_fn_force.py:48-74)."""

    kind = 'linear'

    def __init__(self, zero_intercept, slope, apply_distant_force,
                 apply_nearby_force):
        self.zero_intercept = zero_intercept
        self.slope = slope
        self.apply_distant_force = apply_distant_force
        self.apply_nearby_force = apply_nearby_force
        self.event_horizon = -1. * zero_intercept / slope

    def __call__(self, distance):
        magnitude = self.zero_intercept + self.slope * distance
        if not self.apply_distant_force and distance > self.event_horizon:
            magnitude = 0
        if not self.apply_nearby_force and distance < self.event_horizon:
            magnitude = 0
        return magnitude


class _SpringForceFn(object):
    """Hooke's law (distance_fn_force.py:77-89)."""

    kind = 'spring'

    def __init__(self, spring_constant, equilibrium):
        self.spring_constant = spring_constant
        self.equilibrium = equilibrium

    def __call__(self, distance):
        return -1. * self.spring_constant * (distance - self.equilibrium)


def linear_force_fn(zero_intercept, slope, apply_distant_force=False,
                    apply_nearby_force=True):
    return _LinearForceFn(
        zero_intercept, slope, apply_distant_force, apply_nearby_force)


def spring_force_fn(spring_constant, equilibrium=0):
    return _SpringForceFn(spring_constant, equilibrium)


class DistanceForce(AbstractNewtonianForce):
    """F = force_fn(dist) * unit(diff) on the second sprite
    (distance_fn_force.py:16-45)."""

    def __init__(self, force_fn, symmetric=False):
        self._force_fn = force_fn
        self._symmetric = symmetric


class Tether(AbstractPhysics):
    """Rigidly tethers all sprites of the layers (tether_physics.py:94-140)."""

    def __init__(self, layer_names, update_angle_vel=True, anchor=None):
        super().__init__(updates_per_env_step=1)
        self._layer_names = _as_list(layer_names)
        self._update_angle_vel = update_angle_vel
        self._anchor = anchor


class TetherZippedLayers(AbstractPhysics):
    """Tethers the i-th sprites of each layer (tether_physics.py:143-201)."""

    def __init__(self, layer_names, update_angle_vel=True, anchor=None):
        super().__init__(updates_per_env_step=1)
        self._layer_names = _as_list(layer_names)
        self._update_angle_vel = update_angle_vel
        self._anchor = anchor


class ConstantSpeed(AbstractPhysics):
    """Renormalises velocities to a fixed speed (constant_speed.py:13-46)."""

    def __init__(self, layer_names, speed):
        super().__init__(updates_per_env_step=1)
        self._layer_names = _as_list(layer_names)
        self._speed = speed


class MazePhysics(AbstractPhysics):
    """Constrains the sprites of `avatar_layers` to the grid of the maze whose
    wall sprites are in `maze_layer` (maze_physics.py:19-211).  Must be the last
    corrective physics and needs updates_per_env_step == 1."""

    def __init__(self, maze_layer='walls', avatar_layers=(),
                 constant_speed=None, max_speed=None):
        super().__init__(updates_per_env_step=1)
        self._maze_layer = maze_layer
        self._avatar_layers = avatar_layers
        self._constant_speed = constant_speed
        self._max_speed = max_speed


class RandomMazeWalk(AbstractForce):
    """Walks sprites through the maze at constant speed, turning at random at
    corners and intersections (maze_walk.py:97-196)."""

    def __init__(self, speed, maze_layer='walls', prevent_backtracking=True,
                 allow_wall_backtracking=False, only_turn_at_wall=False):
        self._speed = speed
        self._maze_layer = maze_layer
        self._prevent_backtracking = prevent_backtracking
        self._allow_wall_backtracking = allow_wall_backtracking
        self._only_turn_at_wall = only_turn_at_wall
