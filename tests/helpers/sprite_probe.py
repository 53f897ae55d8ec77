"""Builds a seeded list of sprites with whichever `moog` package is importable in this process
(the repo's MOOG-compatible host package, or the unmodified reference through oracle/shims when
`--reference` is given), mutates some of them the way game rules do (`scale`, `aspect_ratio`,
`angle`, `position`, `shape` assignments; /root/reference/moog/sprite.py:516-558, 616-633) and
prints what the device state is packed from: world vertices, position, circumscribed radius,
moment of inertia, after construction and after every assignment.  JSON on stdout.
Used by tests/test_sprite_vs_reference.py (build container only)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
if '--reference' in sys.argv:
    from oracle import refenv
    refenv.activate()
else:
    import moog_b200  # noqa: F401

import numpy as np  # noqa: E402
from moog import sprite as sprite_lib  # noqa: E402

SHAPES = ['square', 'triangle', 'pentagon', 'hexagon', 'octagon', 'circle', 'star_4', 'star_5', 'star_6',
          'spoke_4', 'spoke_5', 'spoke_6']


def snapshot(s):
    return dict(vertices=np.asarray(s.vertices, dtype=np.float64).tolist(),
                position=[float(s.position[0]), float(s.position[1])],
                max_radius=float(s.max_radius), moment_of_inertia=float(s.moment_of_inertia),
                angle=float(s.angle), scale=float(s.scale), aspect_ratio=float(s.aspect_ratio))


def main():
    rng = np.random.RandomState(123)
    out = []
    for k in range(40):
        if k % 5 == 4:      # custom outlines: clockwise, off-centre, concave (sprite.py:329-394)
            n = rng.randint(3, 9)
            ang = np.sort(rng.uniform(0, 2 * np.pi, n))
            if k % 2:
                ang = ang[::-1]
            shape = (rng.uniform(0.3, 1.0, n)[:, None] * np.stack([np.cos(ang), np.sin(ang)], 1) + rng.uniform(-0.5, 0.5, 2))
        else:
            shape = SHAPES[rng.randint(len(SHAPES))]
        kw = dict(x=rng.uniform(0.1, 0.9), y=rng.uniform(0.1, 0.9), shape=shape, angle=rng.uniform(0, 6.28) * (k % 3 > 0),
                  scale=rng.uniform(0.03, 0.3), aspect_ratio=rng.uniform(0.5, 2.0) if k % 4 else 1.0,
                  mass=rng.uniform(0.5, 3.0), x_vel=rng.uniform(-0.1, 0.1), angle_vel=rng.uniform(-0.1, 0.1))
        s = sprite_lib.Sprite(**kw)
        rec = [snapshot(s)]
        # rule-style assignments; the reference's inertia compounds over scale / aspect changes
        # (sprite.py:411-424) -- two rounds so that a compounding error would show
        for rnd in range(2):
            s.scale = float(rng.uniform(0.03, 0.3))
            rec.append(snapshot(s))
            s.aspect_ratio = float(rng.uniform(0.5, 2.0))
            rec.append(snapshot(s))
            s.angle = float(rng.uniform(0, 6.28))
            rec.append(snapshot(s))
            s.position = np.array([rng.uniform(0.1, 0.9), rng.uniform(0.1, 0.9)])
            rec.append(snapshot(s))
            if rnd == 0 and k % 3 == 0:
                s.shape = SHAPES[rng.randint(len(SHAPES))]
                rec.append(snapshot(s))
        out.append(rec)
    json.dump(out, sys.stdout)


if __name__ == '__main__':
    main()
