"""Sprite generators: callables producing lists of host Sprites by sampling
a factor distribution, with optional rejection against overlap.

Reference: moog/state_initialization/sprite_generators.py:26-190.
"""

import contextlib

import numpy as np

from moog import sprite as sprite_lib

# While a `recording()` block is active every generator call appends a record of what it was
# asked to do and which sprites it produced.  moog_b200.compiler traces a state initializer with
# it to lower the initializer to the device-side reset sampler.
_RECORDS = None


@contextlib.contextmanager
def recording():
    global _RECORDS
    previous, _RECORDS = _RECORDS, []
    try:
        yield _RECORDS
    finally:
        _RECORDS = previous


def generate_sprites(factor_dist, num_sprites=1, max_recursion_depth=int(1e4),
                     fail_gracefully=False):
    """Returns `_generate(disjoint=False, without_overlapping=[])`."""

    def _touches_any(s, others):
        # Every pair is evaluated (no short-circuit), like the reference.
        return any([s.overlaps_sprite(o) for o in others]) if others else False

    def _traced_count():
        """While a state initializer is traced for the device sampler, a random sprite count of the form
        `lambda: np.random.randint(lo, hi)` (functional_maze.py:146) is recorded as the range it is
        drawn from and the trace produces the largest count."""
        calls = []
        real = np.random.randint

        def _randint(low, high=None, size=None, **kwargs):
            if size is not None or kwargs:
                return real(low, high, size, **kwargs)
            lo, hi = (0, low) if high is None else (low, high)
            calls.append((int(lo), int(hi)))
            return int(hi) - 1
        np.random.randint = _randint
        try:
            value = num_sprites()
        finally:
            np.random.randint = real
        if len(calls) == 1 and value == calls[0][1] - 1:
            return value, calls[0]
        return value, None

    def _generate(disjoint=False, without_overlapping=[]):
        count_range = None
        if callable(num_sprites) and _RECORDS is not None:
            count, count_range = _traced_count()
        else:
            count = num_sprites() if callable(num_sprites) else num_sprites
        avoid = list(without_overlapping)
        out = []
        for _ in range(count):
            candidate = sprite_lib.Sprite(**factor_dist.sample())
            rejected = 0
            while _touches_any(candidate, avoid):
                if rejected > max_recursion_depth:
                    if fail_gracefully:
                        return out
                    raise RecursionError(
                        'max_recursion_depth exceeded trying to initialize '
                        'a non-overlapping sprite.')
                rejected += 1
                candidate = sprite_lib.Sprite(**factor_dist.sample())
            out.append(candidate)
            if disjoint:
                avoid = avoid + [candidate]
        if _RECORDS is not None:
            _RECORDS.append(dict(
                factor_dist=factor_dist, num_sprites=num_sprites, count_range=count_range, disjoint=bool(disjoint),
                avoid=list(without_overlapping), out=list(out),
                max_recursion_depth=max_recursion_depth, fail_gracefully=bool(fail_gracefully)))
        return out

    # what moog_b200.compiler lowers a game_rules.CreateSprites(generator=_generate) from
    _generate.factor_dist = factor_dist
    _generate.num_sprites = num_sprites
    _generate.max_recursion_depth = max_recursion_depth
    _generate.fail_gracefully = bool(fail_gracefully)
    return _generate


def chain_generators(*sprite_generators):
    """Concatenates the outputs of several generators."""
    def _generate(*args, **kwargs):
        out = []
        for g in sprite_generators:
            out.extend(g(*args, **kwargs))
        return out
    return _generate


def sample_generator(sprite_generators, p=None):
    """Calls one generator picked at random."""
    def _generate(*args, **kwargs):
        return np.random.choice(sprite_generators, p=p)(*args, **kwargs)
    return _generate


def shuffle(sprite_generator):
    """Randomly permutes a generator's output (z-order randomisation)."""
    def _generate(*args, **kwargs):
        sprites = sprite_generator(*args, **kwargs)
        order = np.arange(len(sprites))
        np.random.shuffle(order)
        return [sprites[i] for i in order]
    return _generate
