"""Known answers for the rarely-hit Collision branches,
recorded by running the UNMODIFIED reference (build container only):

    python oracle/gen_kat_extra.py      # writes tests/golden/kat_rare_branches.json

* crossed bars: two thin rectangles that overlap without any contained vertex
  -> Collision._make_disjoint / _position_correction (collisions.py:586-748)
* C7: the chosen contact is a vertex of sprite_1 inside sprite_0; the
  un-negated perpendicular pushes sprite_0 INTO sprite_1 (collisions.py:268-283)

Each case is one `Physics.step` with updates_per_env_step=1 (forces, then
update_pos_from_vel(1.)); recorded: position and velocity of both sprites.
"""
import collections
import json
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)
from oracle import refenv  # noqa: E402

refenv.activate()
from moog import physics as physics_lib  # noqa: E402
from moog import sprite  # noqa: E402

CASES = {
    'crossed_bars_asymmetric': dict(
        s0=dict(x=.5, y=.5, shape='square', scale=.3, aspect_ratio=.2, x_vel=.01),
        s1=dict(x=.52, y=.5, shape='square', scale=.3, aspect_ratio=.2, angle=float(np.pi / 2)),
        collision=dict(elasticity=1., symmetric=False, update_angle_vel=True)),
    'crossed_bars_symmetric': dict(
        s0=dict(x=.5, y=.5, shape='square', scale=.3, aspect_ratio=.2, x_vel=.01),
        s1=dict(x=.52, y=.5, shape='square', scale=.3, aspect_ratio=.2, angle=float(np.pi / 2)),
        collision=dict(elasticity=1., symmetric=True, update_angle_vel=True)),
    'c7_vertex_of_sprite_1_inside_sprite_0': dict(
        s0=dict(x=.4, y=.5, shape='square', scale=.2, x_vel=.02),
        s1=dict(x=0, y=0, shape=[[.49, .5], [.7, .4], [.7, .6]]),
        collision=dict(elasticity=1., symmetric=False, update_angle_vel=False)),
}


def _sprite(kw):
    kw = dict(kw)
    if not isinstance(kw['shape'], str):
        kw['shape'] = np.array(kw['shape'])
    return sprite.Sprite(**kw)


def main():
    out = {}
    for name, case in CASES.items():
        s0, s1 = _sprite(case['s0']), _sprite(case['s1'])
        force = physics_lib.Collision(**case['collision'])
        # asymmetric entries visit (s0, s1) only; the symmetric one is given one layer
        # per sprite as well, so that the pair order is (s0, s1) in both cases
        state = collections.OrderedDict([('a', [s0]), ('b', [s1])])
        physics = physics_lib.Physics((force, 'a', 'b'), updates_per_env_step=1)
        physics.step(state)
        out[name] = dict(case, expected=dict(
            pos0=[float(v) for v in s0.position], vel0=[float(v) for v in s0.velocity],
            pos1=[float(v) for v in s1.position], vel1=[float(v) for v in s1.velocity],
            still_overlapping=bool(s0.overlaps_sprite(s1))))
        print(name, out[name]['expected'])
    path = os.path.join(_ROOT, 'tests', 'golden', 'kat_rare_branches.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print('->', path)


if __name__ == '__main__':
    main()
