"""Maze inference on the host: `Maze.from_state` restated (reference
moog/maze_lib/maze.py:38-84) for the maze record the device kernels read.

The wall sprites of a maze never move, so the maze matrix the reference infers
at every `Environment.reset()` (maze_physics.py:46-49, maze_walk.py:34-37) is a
function of the initial state; it is evaluated once, when a state is packed,
and stored in the env's `envf` record (include/moog_b200_program.h,
MOOG_MAZE_WORDS).
"""
import numpy as np

_EPSILON = 1e-4        # maze.py:13
_MAX_MAZE_SIZE = 100   # maze.py:17
MAX_MAZE = 32          # MOOG_MAX_MAZE
MAZE_WORDS = 1 + MAX_MAZE


def points_in_closed_path(points, verts):
    """matplotlib `Path.contains_points(points)` (radius 0) for a closed
    polygon: crossing number with the `>=` conventions of _path.h
    point_in_path_impl.  `verts` is the V+1 closed path the
    reference holds (sprite.py:411-424); the duplicated closing vertex adds a
    zero-length edge that never toggles."""
    points = np.asarray(points, dtype=np.float64)
    v = np.asarray(verts, dtype=np.float64)
    n = len(v)
    inside = np.zeros(len(points), dtype=bool)
    if n < 3:
        return inside
    tx, ty = points[:, 0], points[:, 1]
    for i in range(n):
        x0, y0 = v[i]
        x1, y1 = v[(i + 1) % n]
        f0 = y0 >= ty
        f1 = y1 >= ty
        cross = ((y1 - ty) * (x0 - x1) >= (x1 - tx) * (y0 - y1)) == f1
        inside ^= (f0 != f1) & cross
    return inside & np.isfinite(tx) & np.isfinite(ty)


def _contains_points(sp, points):
    """Sprite.contains_points (sprite.py:442-460)."""
    if getattr(sp, 'is_symmetric_circle', False):
        d = points - np.asarray(sp.position, dtype=np.float64)
        return np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) <= sp.max_radius
    path = getattr(sp, '_path', None)
    verts = path.vertices if path is not None else np.concatenate(
        [np.asarray(sp.vertices), np.asarray(sp.vertices)[:1]])
    return points_in_closed_path(points, verts)


def maze_matrix(wall_sprites):
    """-> int array [N, N], maze[j, i] = 1 where the centre of grid cell
    (x index i, y index j) lies inside a wall sprite (maze.py:38-84)."""
    wall_vertices = np.array([np.asarray(s.vertices, dtype=np.float64) for s in wall_sprites])
    maze_size = 1
    while True:
        if maze_size > _MAX_MAZE_SIZE:
            raise ValueError('Cannot find a maze grid size. Your maze sprites are invalid.')
        rounded = np.round(wall_vertices * maze_size) / maze_size
        if np.allclose(rounded, wall_vertices, atol=_EPSILON):
            break
        maze_size += 1
    half = 1. / (2 * maze_size)
    centers = np.linspace(half, 1 - half, maze_size)
    grid = np.stack(np.meshgrid(centers, centers), axis=2)
    flat = np.reshape(grid, (maze_size * maze_size, 2))
    maze = np.zeros(maze_size * maze_size, dtype=bool)
    for s in wall_sprites:
        maze = np.logical_or(maze, _contains_points(s, flat))
    return np.reshape(maze, (maze_size, maze_size)).astype(int)


def maze_record(wall_sprites):
    """-> float64[MAZE_WORDS]: [N, row_0, ..., row_{N-1}, 0...], row j =
    sum(maze[j, i] << i)."""
    m = maze_matrix(wall_sprites)
    n = m.shape[0]
    if n > MAX_MAZE:
        raise ValueError('maze size {} exceeds the device limit {}'.format(n, MAX_MAZE))
    rec = np.zeros(MAZE_WORDS)
    rec[0] = n
    for j in range(n):
        rec[1 + j] = float(sum(int(m[j, i]) << i for i in range(n)))
    return rec
