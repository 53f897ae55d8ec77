"""Binary maze matrix <-> wall sprites (reference: moog/maze_lib/maze.py).

Host-side helper of state initializers: a maze is an N x N 0/1 array
(1 = wall, indexed [row = y cell, column = x cell]) over the unit arena.  The
device kernels never see this class; they read the maze record the packer
derives from the wall sprites (moog_b200/host_maze.py).
"""
import numpy as np

from moog import sprite


class Maze(object):
    """maze.py:20-261 (the methods the shipped configs use)."""

    def __init__(self, maze):
        self.maze = maze
        self.maze_size = maze.shape[0]
        self.grid_side = 1. / self.maze_size
        self.half_grid_side = 0.5 * self.grid_side
        self.side_vertices = np.linspace(
            self.half_grid_side, 1. - self.half_grid_side, self.maze_size)

    @classmethod
    def from_state(cls, state, maze_layer='walls'):
        """Infers the maze from the wall sprites of `maze_layer` (maze.py:38-84)."""
        from moog_b200 import host_maze
        return cls(host_maze.maze_matrix(state[maze_layer]))

    def to_sprites(self, **color):
        """One unit-grid square sprite per wall cell, column-major over (x, y)
        (maze.py:86-111); built at x=0, y=0, scale 1 so that the sprites land
        exactly on their absolute vertices."""
        n = self.maze_size
        edges = np.linspace(0., 1., n + 1)
        out = []
        for x in range(n):
            for y in range(n):
                if not self.maze[y, x]:
                    continue
                square = np.array([[edges[x], edges[y]], [edges[x], edges[y + 1]],
                                   [edges[x + 1], edges[y + 1]], [edges[x + 1], edges[y]]])
                out.append(sprite.Sprite(x=0., y=0., shape=square, **color))
        return out

    def open_vertex(self, i, j):
        if i < 0 or j < 0 or i >= self.maze_size or j >= self.maze_size:
            return False
        return not self.maze[j, i]

    def valid_directions(self, i, j):
        return np.array([[self.open_vertex(k, j) for k in (i - 1, i + 1)],
                         [self.open_vertex(i, k) for k in (j - 1, j + 1)]])

    def sample_open_point(self):
        if np.sum(1 - self.maze) == 0:
            raise ValueError('Maze has no open point.')
        candidates = np.argwhere(self.maze == 0)
        return tuple(candidates[np.random.randint(len(candidates))])

    def sample_distinct_open_points(self, num_points):
        if np.sum(1 - self.maze) < num_points:
            raise ValueError('Maze has no open point.')
        candidates = np.argwhere(self.maze == 0)
        inds = np.random.choice(len(candidates), size=num_points, replace=False)
        return [tuple(candidates[i]) for i in inds]

    def add_outer_walls(self):
        self.maze[0, :] = 1
        self.maze[-1, :] = 1
        self.maze[:, 0] = 1
        self.maze[:, -1] = 1
