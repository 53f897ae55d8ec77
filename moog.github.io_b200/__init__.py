"""moog_b200: B200-native batched implementation of MOOG's Environment.step.

Public surface:
    BatchedEnvironment   -- N independent MOOG envs stepped by sm_100a kernels
    compile_config       -- MOOG config dict -> device program
    pack_states          -- host Sprite states -> SoA state records
"""
from moog_b200.compiler import CompileError, Program, compile_config, pack_states  # noqa: F401
from moog_b200.lambdas import LoweringError  # noqa: F401


def __getattr__(name):
    if name == 'BatchedEnvironment':
        from moog_b200.batched_env import BatchedEnvironment
        return BatchedEnvironment
    raise AttributeError(name)
