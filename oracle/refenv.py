"""Import the UNMODIFIED reference (/root/reference) through the shims.

TEST INFRASTRUCTURE ONLY, and only usable in the build container: the GPU box
has no /root/reference.  Used by oracle/gen_golden.py (golden vectors) and by
the container-only differential tests.
"""

import os
import sys

REFERENCE_ROOT = os.environ.get('MOOG_REFERENCE_ROOT', '/root/reference')
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shims')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'moog'))


def activate():
    """Put shims + reference first on sys.path; returns the `moog` package."""
    if not available():
        raise RuntimeError('reference not found at ' + REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # /root/reference is read-only
    for p in (REFERENCE_ROOT, _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    loaded = sys.modules.get('moog')
    if loaded is not None and not getattr(loaded, '__file__', '').startswith(
            REFERENCE_ROOT):
        raise RuntimeError(
            'a different `moog` package is already imported: {}'.format(
                loaded.__file__))
    import moog  # noqa: F401  pylint: disable=import-outside-toplevel
    return moog


class OverlapLog(object):
    """Records every Sprite.overlaps_sprite call of the reference, in order."""

    def __init__(self):
        from moog import sprite as sprite_lib
        self._cls = sprite_lib.Sprite
        self._orig = None
        self.calls = []

    def __enter__(self):
        orig = self._cls.overlaps_sprite
        self._orig = orig
        calls = self.calls

        def _logged(this, other):
            r = orig(this, other)
            calls.append((this, other, bool(r)))
            return r
        self._cls.overlaps_sprite = _logged
        return self

    def __exit__(self, *exc):
        self._cls.overlaps_sprite = self._orig
