"""Build container only (`reference` marker): this repo's host `moog.sprite.Sprite` -- a rewrite
that computes its outline from the factors instead of transforming a cached path -- against the
UNMODIFIED reference's Sprite (/root/reference/moog/sprite.py:261-424, 516-558, 616-633) on 40
seeded sprites (named shapes, clockwise / off-centre / concave custom outlines), after
construction and after rule-style assignments of scale, aspect_ratio, angle, position and shape,
including the reference's compounding inertia (sprite.py:411-424).  The two packages are both
called `moog`, so each side runs in its own process (tests/helpers/sprite_probe.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _probe(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'helpers', 'sprite_probe.py')] + list(args),
                         capture_output=True, text=True, timeout=600, cwd='/tmp',
                         env=dict(os.environ, PYTHONDONTWRITEBYTECODE='1'))
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout)


@pytest.mark.reference
def test_host_sprite_matches_reference_sprite():
    mine, ref = _probe(), _probe('--reference')
    assert len(mine) == len(ref) == 40
    worst = {}
    for k, (ra, rb) in enumerate(zip(mine, ref)):
        assert len(ra) == len(rb)
        for i, (a, b) in enumerate(zip(ra, rb)):
            for key in b:
                xa, xb = np.array(a[key]), np.array(b[key])
                assert xa.shape == xb.shape, (k, i, key)
                err = float(np.abs(xa - xb).max() / max(np.abs(xb).max(), 1e-12))
                worst[key] = max(worst.get(key, 0.0), err)
    # factors, position (incl. the raw-centroid shift of custom outlines), radius and the compounding
    # inertia are the same floating-point expressions: bit-equal.  World vertices are computed from
    # the factors here and incrementally there: a few ulp.
    for key in ('position', 'angle', 'scale', 'aspect_ratio', 'max_radius', 'moment_of_inertia'):
        assert worst[key] <= 1e-15, (key, worst[key])
    assert worst['vertices'] <= 1e-14, worst['vertices']
