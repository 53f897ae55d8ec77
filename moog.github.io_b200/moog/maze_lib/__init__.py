"""Maze utilities for state initializers (reference: moog/maze_lib/)."""
from moog.maze_lib.maze import Maze  # noqa: F401
from moog.maze_lib.maze_generators import generate_random_maze_matrix  # noqa: F401
