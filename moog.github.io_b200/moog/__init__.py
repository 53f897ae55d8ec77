"""MOOG-compatible host API for the B200 batched environment.

Configs written for jazlab/moog.github.io (`from moog import physics, sprite,
tasks, ...; get_config(level) -> dict`) import this package unchanged: module
names, class names, constructor signatures and the private attribute names the
config compiler reads (`_forces`, `_elasticity`, `_layers_0`, ...) follow the
reference (moog/__init__.py and the sub-package __init__ files).  The classes
here are declarative specs: all stepping happens on the GPU inside
`moog_b200.BatchedEnvironment`; none of them steps sprites on the CPU.
"""
