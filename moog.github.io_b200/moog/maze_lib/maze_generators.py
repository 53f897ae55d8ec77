"""Random maze matrices for state initializers (reference contract:
moog/maze_lib/maze_generators.py:96-246).

`generate_random_maze_matrix(size, ambient_size)` returns a 0/1 matrix
(1 = wall) whose open cells form one connected component with no open 2x2
block and no dead end, optionally centred in a larger all-wall matrix.
Host-side only; draws from the global `np.random` like the reference.
"""
import numpy as np

_MAX_TRIES = 1000


def _opens_a_block(grid, i, j):
    """Would opening (i, j) complete an open 2x2 block?"""
    n = grid.shape[0]
    for di in (-1, 0):
        for dj in (-1, 0):
            a, b = i + di, j + dj
            if a < 0 or b < 0 or a + 1 >= n or b + 1 >= n:
                continue
            closed = 0
            for x in (a, a + 1):
                for y in (b, b + 1):
                    if (x, y) != (i, j) and grid[x, y]:
                        closed += 1
            if closed == 0:
                return True
    return False


def _neighbours(n, i, j):
    return [(a, b) for a, b in ((i - 1, j), (i + 1, j), (i, j - 1), (i, j + 1))
            if 0 <= a < n and 0 <= b < n]


def _grow(size):
    grid = np.ones((size, size))
    start = (np.random.randint(size), np.random.randint(size))
    grid[start] = 0
    frontier = _neighbours(size, *start)
    while frontier:
        k = np.random.randint(len(frontier))
        i, j = frontier.pop(k)
        if not grid[i, j] or _opens_a_block(grid, i, j):
            continue
        grid[i, j] = 0
        for nb in _neighbours(size, i, j):
            if grid[nb] and nb not in frontier:
                frontier.append(nb)
    return grid


def _prune_dead_ends(grid):
    n = grid.shape[0]
    changed = True
    while changed:
        changed = False
        for i in range(n):
            for j in range(n):
                if grid[i, j]:
                    continue
                if sum(1 for nb in _neighbours(n, i, j) if not grid[nb]) < 2:
                    grid[i, j] = 1
                    changed = True


def generate_random_maze_matrix(size, ambient_size=None):
    for _ in range(_MAX_TRIES):
        grid = _grow(size)
        _prune_dead_ends(grid)
        if np.sum(1 - grid) > 0:
            break
    else:
        raise ValueError('could not generate a maze of size {}'.format(size))
    if ambient_size is not None and ambient_size > size:
        out = np.ones((ambient_size, ambient_size))
        lo = (ambient_size - size) // 2
        out[lo:lo + size, lo:lo + size] = grid
        grid = out
    return grid
