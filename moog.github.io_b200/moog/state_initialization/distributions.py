"""Compositional distributions over sprite-factor dictionaries.

Same public surface as the reference's moog/state_initialization/
distributions.py (Continuous :78, Discrete :121, Mixture :159, Intersection
:211, Product :267, SetMinus :319, Selection :367, DependentDistribution
:420): `sample(rng=None) -> dict`, `contains(spec) -> bool`, `keys`.
Sampling runs on the host at reset time and draws from `np.random` in the
same order as the reference, so a seeded config produces the same factors.
"""

import abc

import numpy as np

_MAX_TRIES = int(1e5)


def _rng(rng):
    return np.random if rng is None else rng


def _need(spec, key):
    if key not in spec:
        raise KeyError(
            'key {} is not in spec {}, but must be to evaluate '
            'containment.'.format(key, spec))
    return spec[key]


def _same_keys(components):
    keys = components[0].keys
    for c in components[1:]:
        if c.keys != keys:
            raise ValueError(
                'All components must have the same key sets. However '
                'detected key sets {} and {}'.format(keys, c.keys))
    return keys


def _rejection(draw, accept, what):
    for _ in range(_MAX_TRIES):
        candidate = draw()
        if accept(candidate):
            return candidate
    raise ValueError(
        'Maximum number of tried exceeded when trying to sample from '
        '{}.'.format(what))


class AbstractDistribution(abc.ABC):
    @abc.abstractmethod
    def sample(self, rng=None):
        """Returns a dict of factors."""

    @abc.abstractmethod
    def contains(self, spec):
        """Whether `spec` lies in the support."""

    @property
    @abc.abstractmethod
    def keys(self):
        """Set of factor names produced by sample()."""

    def to_str(self, indent):
        return indent * '  ' + '<{}>'.format(type(self).__name__)

    def __str__(self):
        return self.to_str(indent=0)

    def _get_rng(self, rng=None):
        """np.random unless a generator is given (user subclasses call this, e.g. red_green.py's
        RadialVelocity; reference distributions.py:69-71)."""
        return np.random if rng is None else rng


class Continuous(AbstractDistribution):
    """Uniform on [minval, maxval); float32 by default, like the reference."""

    def __init__(self, key, minval, maxval, dtype='float32'):
        self.key = key
        self.minval = minval
        self.maxval = maxval
        self.dtype = dtype

    def sample(self, rng=None):
        value = _rng(rng).uniform(low=self.minval, high=self.maxval)
        return {self.key: np.asarray(value, dtype=self.dtype)}

    def contains(self, spec):
        v = _need(spec, self.key)
        return v >= self.minval and v < self.maxval

    @property
    def keys(self):
        return {self.key}

    def to_str(self, indent):
        return indent * '  ' + '<Continuous: key={}, mival={}, maxval={}, dtype={}>'.format(
            self.key, self.minval, self.maxval, self.dtype)


class Discrete(AbstractDistribution):
    """Categorical over `candidates` (uniform unless `probs`)."""

    def __init__(self, key, candidates, probs=None):
        self.key = key
        self.candidates = candidates
        self.probs = probs

    def sample(self, rng=None):
        pick = _rng(rng).choice(len(self.candidates), p=self.probs)
        return {self.key: self.candidates[pick]}

    def contains(self, spec):
        return _need(spec, self.key) in self.candidates

    @property
    def keys(self):
        return {self.key}

    def to_str(self, indent):
        return indent * '  ' + '<Discrete: key={}, candidates={}, probs={}>'.format(
            self.key, self.candidates, self.probs)


class Mixture(AbstractDistribution):
    """Picks a component with `probs`, then samples it."""

    def __init__(self, components, probs=None):
        self.components = components
        n = len(components)
        self.probs = np.ones(n) / n if probs is None else np.array(probs)
        self._keys = _same_keys(components)

    def sample(self, rng=None):
        rng = _rng(rng)
        pick = rng.choice(len(self.components), p=self.probs)
        return self.components[pick].sample(rng=rng)

    def contains(self, spec):
        return any(c.contains(spec) for c in self.components)

    @property
    def keys(self):
        return self._keys


class Intersection(AbstractDistribution):
    """Samples component `index_for_sampling`, rejects with the others."""

    def __init__(self, components, index_for_sampling=0):
        self.components = components
        self.index_for_sampling = index_for_sampling
        self._keys = _same_keys(components)

    def sample(self, rng=None):
        rng = _rng(rng)
        source = self.components[self.index_for_sampling]
        return _rejection(lambda: source.sample(rng=rng), self.contains, self)

    def contains(self, spec):
        return all(c.contains(spec) for c in self.components)

    @property
    def keys(self):
        return self._keys


class Product(AbstractDistribution):
    """Independent components over disjoint keys, plus constant factors."""

    def __init__(self, components, **constants):
        self.components = list(components) + [
            Discrete(k, [v]) for k, v in constants.items()]
        self._keys = set()
        total = 0
        for c in self.components:
            self._keys |= set(c.keys)
            total += len(c.keys)
        if len(self._keys) < total:
            raise ValueError(
                'All components must have different keys, yet there are {} '
                'overlapping keys.'.format(total - len(self._keys)))

    def sample(self, rng=None):
        rng = _rng(rng)
        out = {}
        for c in self.components:
            out.update(c.sample(rng=rng))
        return out

    def contains(self, spec):
        return all(c.contains(spec) for c in self.components)

    @property
    def keys(self):
        return self._keys


class _Filtered(AbstractDistribution):
    """base restricted by a second distribution over a subset of its keys."""

    _keep_if_inside = True

    def __init__(self, base, other):
        self.base = base
        self._other = other
        self._keys = base.keys
        if not other.keys.issubset(self._keys):
            raise ValueError(
                'Keys {} is not a subset of keys {} of the base '
                'distribution.'.format(other.keys, base.keys))

    def _ok(self, spec):
        return self._other.contains(spec) == self._keep_if_inside

    def sample(self, rng=None):
        rng = _rng(rng)
        return _rejection(lambda: self.base.sample(rng=rng), self._ok, self)

    def contains(self, spec):
        return self.base.contains(spec) and self._ok(spec)

    @property
    def keys(self):
        return self._keys


class SetMinus(_Filtered):
    """base minus hold_out."""

    _keep_if_inside = False

    def __init__(self, base, hold_out):
        super().__init__(base, hold_out)
        self.hold_out = hold_out


class Selection(_Filtered):
    """base restricted to `filtering`."""

    _keep_if_inside = True

    def __init__(self, base, filtering):
        super().__init__(base, filtering)
        self.filtering = filtering


class DependentDistribution(AbstractDistribution):
    """Some factors are a deterministic function of the sampled ones."""

    def __init__(self, independent_distrib, dependent_fn, dependent_fn_keys):
        self._independent_distrib = independent_distrib
        self._dependent_fn = dependent_fn
        self._dependent_fn_keys = dependent_fn_keys
        if not set(independent_distrib.keys).isdisjoint(set(dependent_fn_keys)):
            raise ValueError(
                'independent_distrib keys {} and dependent_fn keys {} are not '
                'disjoint.'.format(independent_distrib.keys, dependent_fn_keys))

    def sample(self, rng=None):
        out = self._independent_distrib.sample(rng=_rng(rng))
        out.update(self._dependent_fn(out))
        return out

    def contains(self, spec):
        ok = self._independent_distrib.contains(spec)
        derived = self._dependent_fn(
            {k: spec[k] for k in self._independent_distrib.keys})
        for k in self._dependent_fn_keys:
            ok &= spec[k] == derived[k]
        return ok

    @property
    def keys(self):
        return self._independent_distrib.keys.union(self._dependent_fn_keys)
