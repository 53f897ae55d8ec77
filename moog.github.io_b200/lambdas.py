"""Lowering of the small Python callables MOOG configs pass around.

MOOG configs hand plain Python functions to rules and tasks: sprite filters
(`lambda s: s.c2 > 0.6`, cleanup.py:196), sprite modifiers (`def _spoil_fruit
(sprite): sprite.c2 = _BAD_VALUE`, cleanup.py:174-175), pair conditions
(`lambda s_0, s_1: s_1.c2 > T`, cleanup.py:143) and state conditions
(`lambda state: all([s.y < 0. for s in state['prey']])`, pong.py:87).  A
device kernel cannot call them, so they are lowered once, at compile time:

  * sprite-level callables are *traced* with symbolic sprites whose attribute
    reads / writes build an expression tree, emitted as a postfix program for
    the device expression VM (MOOG_X_* in include/moog_b200_program.h);
  * state-level conditions are recognised from their source (all / any / len /
    sum over one layer, combined with not / and / or / comparisons) and become
    MOOG_SC_* ops; contact counters built with `get_contact_counter` are read
    from their closure.

Anything that cannot be lowered raises `LoweringError` at construction time --
there is no per-step Python fallback.
"""

import ast
import collections
import contextlib
import inspect
import textwrap

import numpy as np

(X_END, X_CONST, X_ATTR0, X_ATTR1, X_LT, X_LE, X_GT, X_GE, X_EQ, X_NE, X_AND,
 X_OR, X_NOT, X_ADD, X_SUB, X_MUL, X_DIV, X_NEG, X_ABS, X_MOD, X_STORE,
 X_STORE_POS, X_SELECT, X_ENVF, X_STORE_ENVF, X_RULE_NOISE, X_NORM2) = range(27)

# attribute ids from here on read the sprite's numeric `metadata[key]` columns (MOOG_AT_META0)
AT_META0 = 16
MAX_META_KEYS = 8

ATTRS = ('x', 'y', 'x_vel', 'y_vel', 'angle', 'angle_vel', 'mass', 'scale',
         'aspect_ratio', 'c0', 'c1', 'c2', 'opacity')
# `scale` / `aspect_ratio` assignments re-derive the outline from the shape (sprite.py:546-558 ->
# _set_path :411-424, inertia compounding included), `angle` rotates the cached outline (:531-540)
_WRITABLE = ('x_vel', 'y_vel', 'angle_vel', 'mass', 'c0', 'c1', 'c2',
             'opacity', 'scale', 'aspect_ratio', 'angle')
RESHAPING = ('scale', 'aspect_ratio')


class LoweringError(ValueError):
    pass


# ---------------------------------------------------------------------------
# symbolic expressions
# ---------------------------------------------------------------------------

class Sym(object):
    """Node of a traced expression; `code` is its postfix program."""

    __slots__ = ('code',)
    __hash__ = None

    def __init__(self, code):
        self.code = code

    @staticmethod
    def lift(v):
        if isinstance(v, Sym):
            return v
        if isinstance(v, (bool, int, float)) or hasattr(v, '__float__'):
            return Sym([(X_CONST, 0, float(v))])
        raise LoweringError(
            'cannot use a value of type {} in a device expression'.format(
                type(v).__name__))

    def _bin(self, other, op, swap=False):
        a, b = Sym.lift(self), Sym.lift(other)
        if swap:
            a, b = b, a
        return Sym(a.code + b.code + [(op, 0, 0.0)])

    def __lt__(self, o): return self._bin(o, X_LT)
    def __le__(self, o): return self._bin(o, X_LE)
    def __gt__(self, o): return self._bin(o, X_GT)
    def __ge__(self, o): return self._bin(o, X_GE)
    def __eq__(self, o): return self._bin(o, X_EQ)
    def __ne__(self, o): return self._bin(o, X_NE)
    def __add__(self, o): return self._bin(o, X_ADD)
    def __radd__(self, o): return self._bin(o, X_ADD, True)
    def __sub__(self, o): return self._bin(o, X_SUB)
    def __rsub__(self, o): return self._bin(o, X_SUB, True)
    def __mul__(self, o): return self._bin(o, X_MUL)
    def __rmul__(self, o): return self._bin(o, X_MUL, True)
    def __truediv__(self, o): return self._bin(o, X_DIV)
    def __rtruediv__(self, o): return self._bin(o, X_DIV, True)
    def __mod__(self, o): return self._bin(o, X_MOD)
    def __rmod__(self, o): return self._bin(o, X_MOD, True)
    def __and__(self, o): return self._bin(o, X_AND)
    def __rand__(self, o): return self._bin(o, X_AND, True)
    def __or__(self, o): return self._bin(o, X_OR)
    def __ror__(self, o): return self._bin(o, X_OR, True)
    def __neg__(self): return Sym(self.code + [(X_NEG, 0, 0.0)])
    def __abs__(self): return Sym(self.code + [(X_ABS, 0, 0.0)])
    def __invert__(self): return Sym(self.code + [(X_NOT, 0, 0.0)])

    def __bool__(self):
        path = _PATH[0]
        if path is None:
            raise LoweringError(
                'the callable branches on a per-sprite value (if / and / or / not '
                'on a sprite attribute); write the test as a single comparison or '
                'combine comparisons with & and |')
        return path.decide(self)


class _Path(object):
    """One execution path of a callable that branches on traced values (`if sprite.c0 < 128: ...`):
    the first len(forced) decisions are replayed, later ones are taken True and remembered."""

    def __init__(self, forced):
        self.forced = list(forced)
        self.pos = 0
        self.new = []

    def decide(self, cond):
        k = self.pos
        self.pos += 1
        if k < len(self.forced):
            return self.forced[k]
        self.new.append(cond)
        self.forced.append(True)
        return True


_PATH = [None]
_MAX_PATHS = 64


def _select(cond, a, b):
    """`a if cond else b` as one expression (MOOG_X_SELECT; both sides are evaluated, neither has effects)."""
    if not isinstance(a, Sym) and not isinstance(b, Sym) and type(a) is type(b) and a == b:
        return a
    a, b = Sym.lift(a), Sym.lift(b)
    if a.code == b.code:
        return a
    return Sym(cond.code + a.code + b.code + [(X_SELECT, 0, 0.0)])


def _explore_tree(fn, args, prefix=(), budget=None, hooks=None):
    """Like _explore, but the result is the decision TREE itself -- ('test', condition, subtree if true,
    subtree if false) / ('leaf', value, effects) -- for callables whose tests have to run lazily and in
    order (overlap tests: which ones the reference makes depends on the outcome of the earlier ones).
    hooks = (begin, end): begin() before every path is run, end() -> the effects that path recorded."""
    budget = [_MAX_PATHS] if budget is None else budget
    budget[0] -= 1
    if budget[0] < 0:
        raise LoweringError('the callable branches along more than {} paths'.format(_MAX_PATHS))
    previous, _PATH[0] = _PATH[0], _Path(prefix)
    path = _PATH[0]
    try:
        if hooks:
            hooks[0]()
        value = fn(*args)
        effects = hooks[1]() if hooks else []
    finally:
        _PATH[0] = previous
    node = ('leaf', value, effects)
    base = len(prefix)
    for j in range(len(path.new) - 1, -1, -1):
        other = _explore_tree(fn, args, tuple(path.forced[:base + j]) + (False,), budget, hooks)
        node = ('test', path.new[j], node, other)
    return node


def _explore(fn, args, prefix=(), budget=None):
    """Value of fn(*args) over every path its data-dependent `if`s can take, folded into selects."""
    budget = [_MAX_PATHS] if budget is None else budget
    budget[0] -= 1
    if budget[0] < 0:
        raise LoweringError('the callable branches on per-sprite values along more than {} paths'.format(_MAX_PATHS))
    previous, _PATH[0] = _PATH[0], _Path(prefix)
    path = _PATH[0]
    try:
        value = fn(*args)
    finally:
        _PATH[0] = previous
    base = len(prefix)
    for j in range(len(path.new) - 1, -1, -1):
        other = _explore(fn, args, tuple(path.forced[:base + j]) + (False,), budget)
        value = _select(path.new[j], value, other)
    return value


# `sprite.metadata[key]` of numeric (or bool) values: the keys a program's callables read, in the order
# they were first seen; column k lives in the state record (envf block MOOG_H_META_OFF) and is read
# as attribute AT_META0 + k.  compiler.compile_config installs the list for the program it builds.
_META_KEYS = [None]


@contextlib.contextmanager
def metadata_columns(keys):
    previous, _META_KEYS[0] = _META_KEYS[0], keys
    try:
        yield keys
    finally:
        _META_KEYS[0] = previous


def _attr_read(which, attr):
    """Code entry reading attribute `attr` of sprite 0 / sprite 1 of a pair callable, or -- `which` =
    ('bound', k) -- of the k-th sprite a state-level callable picked out of the state (`state[layer][i]`):
    the binding travels in the high bits of the argument until _bind_pair assigns it to sprite 0 or 1."""
    if isinstance(which, tuple):
        return (X_ATTR0, attr | ((which[1] + 1) << 8), 0.0)
    return (X_ATTR0 if which == 0 else X_ATTR1, attr, 0.0)


class SymOverlap(object):
    """`a.overlaps_sprite(b)` between two sprites picked out of the state: only usable as a test."""

    def __init__(self, a, b):
        if not (isinstance(a, SymSprite) and isinstance(b, SymSprite) and isinstance(a._which, tuple)
                and isinstance(b._which, tuple)):
            raise LoweringError('overlaps_sprite is lowered between two sprites taken from the state only')
        self.pair = (a._which[1], b._which[1])
        self.owner = getattr(a, '_owner', None)

    def decide(self):
        """The test is made where the CALL is made (`[a.overlaps_sprite(s) for s in layer]` calls for every
        sprite before any() looks at the results): the outcome is a plain bool on this path."""
        path = _PATH[0]
        if path is None:
            raise LoweringError('overlaps_sprite between sprites of the state is lowered in state-level callables only')
        owner = self.owner
        if owner is not None and owner.geometry_touched():
            raise LoweringError('an overlap test after the callable moved / reshaped a sprite is not lowered')
        return path.decide(self)


class SymHas(object):
    """`index < len(state[layer])` while a state-level callable iterates over a layer."""

    def __init__(self, layer, index):
        self.layer, self.index = layer, index


class SymMetadata(object):
    """`sprite.metadata` while tracing: item reads become reads of a metadata column."""

    def __init__(self, which):
        self._which = which

    def __getitem__(self, key):
        keys = _META_KEYS[0]
        if keys is None:
            raise LoweringError('sprite.metadata is not available to this callable on the device')
        if not isinstance(key, str):
            raise LoweringError('sprite.metadata keys must be strings on the device path')
        if key not in keys:
            if len(keys) >= MAX_META_KEYS:
                raise LoweringError('at most {} metadata keys are carried on the device'.format(MAX_META_KEYS))
            keys.append(key)
        return Sym([_attr_read(self._which, AT_META0 + keys.index(key))])

    def get(self, key, default=None):
        raise LoweringError('sprite.metadata.get() is not lowered: read the key directly')


class SymSprite(object):
    """Stand-in for a Sprite while tracing: reads give Sym nodes, writes are
    recorded as stores."""

    def __init__(self, which, stores=None):
        object.__setattr__(self, '_which', which)
        object.__setattr__(self, '_stores', stores)

    def __getattr__(self, name):
        if name in ATTRS:
            return Sym([_attr_read(self._which, ATTRS.index(name))])
        if name == 'overlaps_sprite' and isinstance(self._which, tuple):
            return lambda other: SymOverlap(self, other).decide()
        if name == 'position':
            return SymVec([self.x, self.y])
        if name == 'velocity':
            return SymVec([self.x_vel, self.y_vel])
        if name == 'metadata':
            return SymMetadata(self._which)
        raise LoweringError(
            'sprite attribute {!r} is not available on the device'.format(name))

    def __setattr__(self, name, value):
        if self._stores is None:
            raise LoweringError('this callable may not modify sprites')
        if name == 'position':
            self._stores.append(('position', (Sym.lift(value[0]), Sym.lift(value[1]))))
            return
        if name == 'velocity':
            self._stores.append(('velocity', (Sym.lift(value[0]), Sym.lift(value[1]))))
            return
        if name not in _WRITABLE:
            raise LoweringError(
                'assigning sprite.{} is not supported on the device path '
                '(writable: {})'.format(name, ', '.join(_WRITABLE)))
        self._stores.append((name, Sym.lift(value)))


class SymVec(object):
    """A traced 2-vector (`sprite.velocity`, `sprite.position`): arithmetic and comparisons are
    elementwise (with scalars, sequences of two numbers or other SymVecs), NumPy ufuncs applied
    to it are mapped to the same operators, `np.all` / `np.any` (and, after the AST rewrite of
    `_symbolic_call`, the builtins `all` / `any`) reduce with & / |."""

    __array_priority__ = 1000
    __hash__ = None

    def __init__(self, elems):
        self.elems = list(elems)

    def _zip(self, other):
        if isinstance(other, SymVec):
            others = other.elems
        elif hasattr(other, '__len__') and not isinstance(other, (str, bytes)):
            others = list(other)
            if len(others) != len(self.elems):
                raise LoweringError('cannot combine a sprite 2-vector with a sequence of {} values'.format(len(others)))
        else:
            others = [other] * len(self.elems)
        return zip(self.elems, others)

    def _map(self, other, fn):
        return SymVec([fn(Sym.lift(a), b) for a, b in self._zip(other)])

    def __eq__(self, o): return self._map(o, lambda a, b: a == b)
    def __ne__(self, o): return self._map(o, lambda a, b: a != b)
    def __lt__(self, o): return self._map(o, lambda a, b: a < b)
    def __le__(self, o): return self._map(o, lambda a, b: a <= b)
    def __gt__(self, o): return self._map(o, lambda a, b: a > b)
    def __ge__(self, o): return self._map(o, lambda a, b: a >= b)
    def __add__(self, o): return self._map(o, lambda a, b: a + b)
    def __radd__(self, o): return self._map(o, lambda a, b: b + a)
    def __sub__(self, o): return self._map(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._map(o, lambda a, b: b - a)
    def __mul__(self, o): return self._map(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._map(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._map(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._map(o, lambda a, b: b / a)
    def __mod__(self, o): return self._map(o, lambda a, b: a % b)
    def __and__(self, o): return self._map(o, lambda a, b: a & b)
    def __or__(self, o): return self._map(o, lambda a, b: a | b)
    def __neg__(self): return SymVec([-Sym.lift(a) for a in self.elems])
    def __abs__(self): return SymVec([abs(Sym.lift(a)) for a in self.elems])
    def __invert__(self): return SymVec([~Sym.lift(a) for a in self.elems])

    _UFUNCS = {'add': '__add__', 'subtract': '__sub__', 'multiply': '__mul__', 'true_divide': '__truediv__',
               'divide': '__truediv__', 'remainder': '__mod__', 'mod': '__mod__', 'less': '__lt__',
               'less_equal': '__le__', 'greater': '__gt__', 'greater_equal': '__ge__', 'equal': '__eq__',
               'not_equal': '__ne__', 'logical_and': '__and__', 'logical_or': '__or__',
               'bitwise_and': '__and__', 'bitwise_or': '__or__'}
    _UNARY = {'negative': '__neg__', 'absolute': '__abs__', 'logical_not': '__invert__'}
    _REVERSED = {'__add__': '__radd__', '__sub__': '__rsub__', '__mul__': '__rmul__', '__truediv__': '__rtruediv__'}

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != '__call__' or kwargs:
            return NotImplemented
        name = ufunc.__name__
        if name in self._UNARY and len(inputs) == 1:
            return getattr(self, self._UNARY[name])()
        if name in self._UFUNCS and len(inputs) == 2:
            op = self._UFUNCS[name]
            if inputs[0] is self:
                return getattr(self, op)(inputs[1])
            if op in self._REVERSED:
                return getattr(self, self._REVERSED[op])(inputs[0])
            return getattr(SymVec(list(inputs[0]) if hasattr(inputs[0], '__len__') else [inputs[0]] * len(self.elems)),
                           op)(self)
        if name == 'matmul' and len(inputs) == 2 and inputs[1] is self and len(self.elems) == 2:
            # a 2 x 2 matrix of 0 / +-1 entries times the vector (a quarter turn, a reflection): every product is
            # exact, so the order BLAS adds them in cannot matter
            m = np.asarray(inputs[0])
            if m.shape == (2, 2) and m.dtype != object and np.isin(m, (-1, 0, 1)).all():
                x, y = (Sym.lift(v) for v in self.elems)
                return SymVec([float(m[0, 0]) * x + float(m[0, 1]) * y, float(m[1, 0]) * x + float(m[1, 1]) * y])
        raise LoweringError('numpy.{} of a sprite vector is not available on the device'.format(name))

    def __array_function__(self, func, types, args, kwargs):
        if func is np.linalg.norm and len(args) == 1 and args[0] is self and not kwargs and len(self.elems) == 2:
            x, y = (Sym.lift(v) for v in self.elems)       # np.linalg.norm of a 2-vector: sqrt(dot(v, v))
            return Sym(x.code + y.code + [(X_NORM2, 0, 0.0)])
        if func is np.all:
            return self.all()
        if func is np.any:
            return self.any()
        impl = getattr(func, '_implementation', None)     # anything else: NumPy's own code, as without this protocol
        if impl is None:
            raise LoweringError('numpy.{} of a sprite vector is not available on the device'.format(getattr(func, '__name__', func)))
        return impl(*args, **kwargs)

    def __len__(self):
        return len(self.elems)

    def __iter__(self):
        return iter(self.elems)

    def __getitem__(self, k):
        return self.elems[k]

    def all(self, *_, **__):       # np.all(vec) dispatches here
        out = Sym.lift(self.elems[0])
        for x in self.elems[1:]:
            out = out & x
        return out

    def any(self, *_, **__):       # np.any(vec)
        out = Sym.lift(self.elems[0])
        for x in self.elems[1:]:
            out = out | x
        return out


_SymFirstSprite = SymSprite   # `state[layer][0]` while tracing a state condition


class _SymLayer(object):
    def __init__(self, owner, name):
        self._owner, self._name = owner, name

    def __getitem__(self, k):
        if k != 0:
            raise LoweringError('only state[layer][0] can be read in a state condition')
        self._owner.layers.append(self._name)
        return _SymFirstSprite(0)


class SymState(object):
    """Stand-in for the environment state while tracing `state[layer][0].attr`."""

    def __init__(self):
        self.layers = []

    def __getitem__(self, name):
        return _SymLayer(self, name)


class SymMetaValue(Sym):
    """`meta_state[key]` while tracing: a variable of the env (envf slot); strings compare by their codes."""

    __slots__ = ('prog',)
    __hash__ = None

    def __init__(self, code, prog):
        Sym.__init__(self, code)
        self.prog = prog

    def _bin(self, other, op, swap=False):
        if isinstance(other, str):
            other = self.prog.intern(other)
        return Sym._bin(self, other, op, swap)


class SymMetaState(object):
    """The env's `meta_state` dict while a state-level callable is traced: reads are reads of the env's
    variables; writes (only where effects are allowed) are recorded as effects of the path."""

    def __init__(self, prog, owner=None):
        self._prog, self._owner = prog, owner

    def __getitem__(self, key):
        if self._owner is not None and ('meta', key) in self._owner.shadow:
            return self._owner.shadow[('meta', key)]
        return SymMetaValue([(X_ENVF, self._prog.meta_slot(key), 0.0)], self._prog)

    def __setitem__(self, key, value):
        owner = self._owner
        if owner is None or owner.effects is None:
            raise LoweringError('this callable may not modify meta_state')
        if isinstance(value, str):
            value = self._prog.intern(value)
        value = Sym.lift(value)
        owner.shadow[('meta', key)] = value
        owner.effects.append(('var', self._prog.meta_slot(key), value))

    def __contains__(self, key):
        return key in self._prog.meta_vars

    def get(self, key, default=None):
        return self[key] if key in self._prog.meta_vars else default


def randint_range(fn):
    """`fn` of the form `lambda: np.random.randint(lo, hi)` -> (lo, hi); any other callable -> None.
    (Phase durations, task_phases.py:60-63; sprite counts, sprite_generators.py.)"""
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
        value = fn()
    except Exception:  # pylint: disable=broad-except
        return None
    finally:
        np.random.randint = real
    if len(calls) == 1 and value == calls[0][1] - 1 and calls[0][1] > calls[0][0]:
        return calls[0]
    return None


class BoundSprite(SymSprite):
    """A sprite picked out of the state by a state-level callable.  Reads see what the callable itself
    assigned earlier on this path; assignments are recorded as effects of the path (a rule's `step`)."""

    def __init__(self, owner, k):
        SymSprite.__init__(self, ('bound', k), None)
        object.__setattr__(self, '_owner', owner)
        object.__setattr__(self, '_k', k)

    def __getattr__(self, name):
        shadow = self._owner.shadow
        if (self._k, name) in shadow:
            return shadow[(self._k, name)]
        if name == 'velocity' and ((self._k, 'x_vel') in shadow or (self._k, 'y_vel') in shadow):
            return SymVec([self.x_vel, self.y_vel])
        if name == 'position' and ((self._k, 'x') in shadow or (self._k, 'y') in shadow):
            return SymVec([self.x, self.y])
        return SymSprite.__getattr__(self, name)

    def __setattr__(self, name, value):
        owner = self._owner
        if owner.effects is None:
            raise LoweringError('this callable may not modify sprites')
        if name in ('position', 'velocity'):
            pair = (Sym.lift(value[0]), Sym.lift(value[1]))
            a, b = ('x', 'y') if name == 'position' else ('x_vel', 'y_vel')
            owner.shadow[(self._k, a)], owner.shadow[(self._k, b)] = pair
            owner.effects.append(('sprite', self._k, name, pair))
            return
        if name not in _WRITABLE:
            raise LoweringError('assigning sprite.{} is not supported on the device path (writable: {})'.format(
                name, ', '.join(_WRITABLE)))
        value = Sym.lift(value)
        owner.shadow[(self._k, name)] = value
        owner.effects.append(('sprite', self._k, name, value))


class _BoundLayer(object):
    def __init__(self, owner, name):
        self._owner, self._name = owner, name

    def _sprite(self, index):
        key = (self._name, index)
        if key not in self._owner.bound:
            self._owner.bound.append(key)
        return BoundSprite(self._owner, self._owner.bound.index(key))

    def __getitem__(self, index):
        if not isinstance(index, int) or index < 0:
            raise LoweringError('state[{!r}][...] must be a fixed non-negative index in a state-level callable'.format(self._name))
        return self._sprite(index)

    def __iter__(self):
        """`for s in state[layer]`: one path per number of sprites the layer can hold (a test of
        `index < len(layer)` before each one)."""
        path, cap = _PATH[0], self._owner.capacity(self._name)
        if path is None or cap is None:
            raise LoweringError('iterating over state[{!r}] is not lowered here'.format(self._name))
        for index in range(cap):
            if not path.decide(SymHas(self._name, index)):
                return
            yield self._sprite(index)

    def __len__(self):
        raise LoweringError('len(state[{!r}]) is not lowered here'.format(self._name))


class BoundState(object):
    """The environment state for a state-level callable that picks single sprites (`state['agent'][0]`),
    reads their attributes / metadata and tests overlaps between them (bounce_box_contact_prediction.py:94-103)."""

    _GEOMETRY = ('x', 'y', 'position', 'angle', 'scale', 'aspect_ratio')

    def __init__(self, prog=None, effects=False):
        self.bound = []       # (layer name, index) of every sprite picked, in first-use order (all paths)
        self.prog = prog
        self.shadow = {}      # this path: (bound id, attribute) -> value the callable assigned
        self.effects = [] if effects else None

    def begin_path(self):
        self.shadow = {}
        if self.effects is not None:
            self.effects = []

    def capacity(self, layer):
        if self.prog is None:
            return None
        return self.prog.layer_cap[self.prog.layer_index(layer)]

    def geometry_touched(self):
        return any(e[0] == 'sprite' and e[2] in self._GEOMETRY for e in (self.effects or ()))

    def __getitem__(self, name):
        return _BoundLayer(self, name)


def _bind_pair(code, what, first=None):
    """Expression over bound sprites -> (code over sprite 0 / sprite 1, [bound ids of sprite 0, sprite 1]);
    `first`: the bound sprite that must be sprite 0 (the target of the stores in `code`)."""
    ids = [] if first is None else [first]
    for op, arg, _ in code:
        if op in (X_ATTR0, X_ATTR1) and arg >> 8:
            k = (arg >> 8) - 1
            if k not in ids:
                ids.append(k)
    if len(ids) > 2:
        raise LoweringError('{} reads more than two sprites of the state in one expression'.format(what))
    out = []
    for op, arg, c in code:
        if op in (X_ATTR0, X_ATTR1) and arg >> 8:
            out.append((X_ATTR0 if ids.index((arg >> 8) - 1) == 0 else X_ATTR1, arg & 0xff, c))
        else:
            out.append((op, arg, c))
    return out, ids


# node kinds of a decision tree (MOOG_SC_TREE / MOOG_R_TREE)
TN_LEAF, TN_TEST_EXPR, TN_TEST_OVERLAP, TN_TEST_HAS, TN_ACTION = range(5)


def _emit_tree(tree, state, prog, what, with_value):
    """Decision tree -> ipool nodes of 8 ints (kind, expr, layer / index of sprite 0, layer / index of
    sprite 1, next if true, next if false); returns (ipool start, number of nodes)."""
    nodes = []

    def sprite_ref(k):
        if k is None:
            return (-1, 0)
        layer, index = state.bound[k]
        return (prog.layer_index(layer), index)

    def refs(ids):
        ids = list(ids) + [None] * (2 - len(ids))
        return sprite_ref(ids[0]) + sprite_ref(ids[1])

    def effect_code(effect):
        """One recorded assignment -> (code ending in its store, bound id of the sprite stored to or None)."""
        if effect[0] == 'var':
            return Sym.lift(effect[2]).code + [(X_STORE_ENVF, effect[1], 0.0)], None
        _, k, name, value = effect
        return _stores_to_code([(name, value)])[:-1], k

    def emit(node):
        me = len(nodes)
        nodes.append(None)
        if node[0] == 'leaf':
            effects = list(node[2])
            written = set()
            chain = []
            for effect in effects:
                code, target = effect_code(effect)
                reads = {('sprite', (arg >> 8) - 1, arg & 0xff) for op, arg, _ in code if op in (X_ATTR0, X_ATTR1) and arg >> 8}
                reads |= {('var', arg) for op, arg, _ in code if op == X_ENVF}
                if reads & written:
                    raise LoweringError('{}: an assignment reads a value an earlier assignment of the same call changed '
                                        '(effects are applied one after the other on the device)'.format(what))
                if effect[0] == 'var':
                    written.add(('var', effect[1]))
                else:
                    for attr in ({'position': ('x', 'y'), 'velocity': ('x_vel', 'y_vel')}.get(effect[2], (effect[2],))):
                        written.add(('sprite', effect[1], ATTRS.index(attr)))
                bound_code, ids = _bind_pair(code, what, first=target)
                chain.append((TN_ACTION, prog.add_expr(bound_code + [(X_CONST, 0, 1.0)])) + refs(ids))
            if with_value:
                value = node[1]
                code, ids = _bind_pair(Sym.lift(value).code, what)
                last = (TN_LEAF, prog.add_expr(code)) + refs(ids) + (0, 0)
            else:
                last = (TN_LEAF, -1, -1, 0, -1, 0, 0, 0)
            # the actions of the path, then the leaf
            slots = [me] + [None] * len(chain)
            for q in range(len(chain)):
                slots[q + 1] = len(nodes)
                nodes.append(None)
            for q, head in enumerate(chain):
                nodes[slots[q]] = head + (slots[q + 1], slots[q + 1])
            nodes[slots[-1]] = last
            return me
        cond = node[1]
        if isinstance(cond, SymOverlap):
            head = (TN_TEST_OVERLAP, -1) + refs(cond.pair)
        elif isinstance(cond, SymHas):
            head = (TN_TEST_HAS, -1, prog.layer_index(cond.layer), cond.index, -1, 0)
        else:
            code, ids = _bind_pair(cond.code, what)
            head = (TN_TEST_EXPR, prog.add_expr(code)) + refs(ids)
        yes = emit(node[2])
        no = emit(node[3])
        nodes[me] = head + (yes, no)
        return me

    emit(tree)
    flat = [v for nd in nodes for v in nd]
    return prog.add_ints(flat), len(nodes)


def state_tree(fn, prog, what):
    """A state-level callable `fn(state[, meta_state])` made of single-sprite picks, attribute / metadata /
    meta_state reads, overlap tests and `if`s -> index of a MOOG_SC_TREE op: a decision tree evaluated
    lazily, test by test, in the order Python would make them (so the overlap calls are the reference's,
    call for call)."""
    state = BoundState(prog)
    call = fn
    try:
        call = _rewritten(fn)
    except Exception:  # pylint: disable=broad-except
        call = fn
    try:
        arity = _n_params(fn)
    except (TypeError, ValueError):
        arity = 1
    args = (state,) if arity < 2 else (state, SymMetaState(prog))
    with no_randomness(what):
        try:
            tree = _explore_tree(call, args, hooks=(state.begin_path, lambda: []))
        except LoweringError:
            raise
        except Exception as exc:  # pylint: disable=broad-except
            raise LoweringError('{} cannot be lowered to a decision tree over the state ({}: {})'.format(
                what, type(exc).__name__, exc))
    start, count = _emit_tree(tree, state, prog, what, with_value=True)
    return prog.emit(170, 0, (start, count))   # MOOG_SC_TREE


class _RuleProxy(object):
    """`self` of a user-defined rule while its reset / step is traced: numeric attributes the rule assigns
    are its state variables (one envf slot each); everything else reads through to the real object;
    methods are re-bound to the proxy."""

    def __init__(self, rule, variables):
        object.__setattr__(self, '_rule', rule)
        object.__setattr__(self, '_values', dict(variables))
        object.__setattr__(self, '_assigned', [])

    def __getattr__(self, name):
        values = object.__getattribute__(self, '_values')
        if name in values:
            return values[name]
        value = getattr(object.__getattribute__(self, '_rule'), name)
        if inspect.ismethod(value):
            return value.__func__.__get__(self)
        return value

    def __setattr__(self, name, value):
        object.__getattribute__(self, '_values')[name] = value
        object.__getattribute__(self, '_assigned').append(name)


def trace_rule(rule, prog):
    """A user-defined rule class (functional_maze.py:18-67 `Booster`: a countdown of its own, an overlap
    test over a layer, attribute changes on a sprite) -> (nodes start, node count, envf slot of its first
    state variable, initial values): `reset` is traced for the variables and their initial values, `step`
    along every path into a decision tree whose leaves carry the assignments of that path
    (MOOG_R_TREE).  What it may do: pick sprites (`state[layer][i]`, `for s in state[layer]`), read their
    attributes / metadata, test overlaps, branch, assign writable sprite attributes and its own numeric
    attributes.  Anything else is refused."""
    what = 'rule {}'.format(type(rule).__name__)
    cls = type(rule)

    def numeric(v):
        return isinstance(v, (bool, int, float)) or (hasattr(v, '__float__') and not isinstance(v, (Sym, SymVec)))

    # reset(): which attributes are state, and what they start from
    proxy = _RuleProxy(rule, {})
    with no_randomness(what):
        try:
            cls.reset(proxy, BoundState(prog), None)
        except LoweringError:
            raise
        except Exception as exc:  # pylint: disable=broad-except
            raise LoweringError('{}.reset cannot be traced ({}: {})'.format(what, type(exc).__name__, exc))
    initial = collections.OrderedDict()
    for name in proxy._assigned:
        v = proxy._values[name]
        if not numeric(v):
            raise LoweringError('{}.reset assigns self.{} = {!r}: only numbers are carried on the device'.format(what, name, v))
        initial[name] = float(v)

    draws = []          # (kind, rule-noise column, parameter) of the k-th random draw step() makes, on every path

    def draw(kind, k, param):
        if k == len(draws):
            draws.append((kind, prog.rule_noise_dim, param))
            prog.rule_noise_dim += 1
        if draws[k][0] != kind or draws[k][2] != param:
            raise LoweringError('{}.step draws random numbers in an order that depends on the state'.format(what))
        return Sym([(X_RULE_NOISE, draws[k][1], 0.0)])

    @contextlib.contextmanager
    def traced_draws(counter):
        """np.random.uniform(lo, hi) and np.random.randint(n) (scalars) become reads of rule-noise columns:
        lo + (hi - lo) * u, and the number of thresholds k / n that u has passed."""
        saved = (np.random.uniform, np.random.randint)

        def uniform(low=0.0, high=1.0, size=None):
            if size is not None:
                raise LoweringError('{}.step draws an array of random numbers'.format(what))
            u = draw('uniform', counter[0], None)
            counter[0] += 1
            return low + (high - low) * u

        def randint(low, high=None, size=None, **kwargs):
            if size is not None or kwargs:
                raise LoweringError('{}.step draws an array of random integers'.format(what))
            lo, hi = (0, low) if high is None else (low, high)
            n = int(hi) - int(lo)
            if n < 1 or n > 16:
                raise LoweringError('{}.step: np.random.randint over {} values is not lowered'.format(what, n))
            u = draw('randint', counter[0], n)
            counter[0] += 1
            out = Sym.lift(float(lo))
            for k in range(1, n):
                out = out + (u >= (k / n))
            return out
        np.random.uniform, np.random.randint = uniform, randint
        try:
            yield
        finally:
            np.random.uniform, np.random.randint = saved

    def explore(variables, slots):
        state = BoundState(prog, effects=True)
        holder = {}
        counter = [0]

        def begin():
            state.begin_path()
            counter[0] = 0
            holder['proxy'] = _RuleProxy(rule, {n: Sym([(X_ENVF, slots[n], 0.0)]) for n in variables})

        def end():
            px = holder['proxy']
            effects = list(state.effects)
            for name in dict.fromkeys(px._assigned):
                v = px._values[name]
                if isinstance(v, SymVec) or not (isinstance(v, Sym) or numeric(v)):
                    raise LoweringError('{}.step assigns self.{} = {!r}: only numbers are carried on the device'.format(
                        what, name, v))
                effects.append(('var', slots.get(name, -1), v, name))
            return effects

        def call(st):
            out = cls.step(holder['proxy'], st, SymMetaState(prog, st))
            if out is not None:
                raise LoweringError('{}.step returns a value'.format(what))

        with no_randomness(what), traced_draws(counter):
            try:
                tree = _explore_tree(call, (state,), hooks=(begin, end))
            except LoweringError:
                raise
            except Exception as exc:  # pylint: disable=broad-except
                raise LoweringError('{}.step cannot be lowered to a decision tree over the state ({}: {})'.format(
                    what, type(exc).__name__, exc))
        return tree, state

    def assigned_names(tree, out):
        if tree[0] == 'leaf':
            for e in tree[2]:
                if e[0] == 'var' and len(e) > 3 and e[3] not in out:
                    out.append(e[3])
        else:
            assigned_names(tree[2], out)
            assigned_names(tree[3], out)
        return out

    # first pass: discover the variables step() assigns that reset() did not (they start from the object's own value)
    probe_slots = {n: k for k, n in enumerate(initial)}
    tree, _ = explore(list(initial), probe_slots)
    for name in assigned_names(tree, []):
        if name not in initial:
            v = getattr(rule, name, None)
            if not numeric(v):
                raise LoweringError('{}.step assigns self.{}, which has no numeric value before the first step'.format(what, name))
            initial[name] = float(v)
    base = prog.alloc_envf(max(len(initial), 1))
    slots = {n: base + k for k, n in enumerate(initial)}
    tree, state = explore(list(initial), slots)

    def strip(node):        # ('var', slot, value, name) -> ('var', slot, value)
        if node[0] == 'leaf':
            return ('leaf', node[1], [e[:3] if e[0] == 'var' else e for e in node[2]])
        return ('test', node[1], strip(node[2]), strip(node[3]))

    start, count = _emit_tree(strip(tree), state, prog, what, with_value=False)
    if draws:
        prog.rule_draws.append((rule, list(draws)))
    return start, count, base, list(initial.values())


def _n_params(fn):
    return len(inspect.signature(fn).parameters)


def closure_vars(fn):
    """{free variable name: value} of a Python function (empty if none)."""
    code = getattr(fn, '__code__', None)
    cells = getattr(fn, '__closure__', None)
    if code is None or not cells:
        return {}
    out = {}
    for name, cell in zip(code.co_freevars, cells):
        try:
            out[name] = cell.cell_contents
        except ValueError:
            pass
    return out


def _unwrap(fn, name):
    """Peel the reference's signature-adapting lambdas, e.g.
    `lambda s_a, s_t, meta_state: condition(s_a, s_t)` (contact_reward.py:60)."""
    seen = 0
    while callable(fn) and seen < 4:
        cv = closure_vars(fn)
        if set(cv) == {name} and getattr(fn, '__name__', '') == '<lambda>':
            fn = cv[name]
            seen += 1
        else:
            break
    return fn


def _code_of(value):
    return Sym.lift(value).code


# ---------------------------------------------------------------------------
# sprite-level callables
# ---------------------------------------------------------------------------

def _sym_any(x):
    if isinstance(x, SymVec):
        return x.any()
    items = list(x)
    if any(isinstance(v, (Sym, SymVec)) for v in items):
        out = Sym.lift(items[0].any() if isinstance(items[0], SymVec) else items[0])
        for v in items[1:]:
            out = out | (v.any() if isinstance(v, SymVec) else v)
        return out
    return any(items)


def _sym_all(x):
    if isinstance(x, SymVec):
        return x.all()
    items = list(x)
    if any(isinstance(v, (Sym, SymVec)) for v in items):
        out = Sym.lift(items[0].all() if isinstance(items[0], SymVec) else items[0])
        for v in items[1:]:
            out = out & (v.all() if isinstance(v, SymVec) else v)
        return out
    return all(items)


def _sym_not(x):
    return ~x if isinstance(x, (Sym, SymVec)) else (not x)


def _sym_and(a, b):
    return a & b if isinstance(a, (Sym, SymVec)) or isinstance(b, (Sym, SymVec)) else (a and b)


def _sym_or(a, b):
    return a | b if isinstance(a, (Sym, SymVec)) or isinstance(b, (Sym, SymVec)) else (a or b)


class _BoolRewriter(ast.NodeTransformer):
    """`a and b` / `a or b` / `not a` / `any(v)` / `all(v)` -> calls that build expressions when an
    operand is symbolic.  (Both operands of and / or are evaluated: a per-sprite expression has no
    side effect to short-circuit.)"""

    def visit_BoolOp(self, node):
        self.generic_visit(node)
        fn = '_moog_and' if isinstance(node.op, ast.And) else '_moog_or'
        out = node.values[0]
        for v in node.values[1:]:
            out = ast.Call(func=ast.Name(id=fn, ctx=ast.Load()), args=[out, v], keywords=[])
        return out

    def visit_UnaryOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Not):
            return ast.Call(func=ast.Name(id='_moog_not', ctx=ast.Load()), args=[node.operand], keywords=[])
        return node

    def visit_Call(self, node):
        self.generic_visit(node)
        if isinstance(node.func, ast.Name) and node.func.id in ('any', 'all') and len(node.args) == 1:
            node.func = ast.Name(id='_moog_' + node.func.id, ctx=ast.Load())
        return node


def _rewritten(fn):
    """`fn` recompiled with _BoolRewriter applied (same globals and closure values)."""
    node = _lambda_ast(fn)
    node = _BoolRewriter().visit(node)
    ns = dict(getattr(fn, '__globals__', {}))
    ns.update(closure_vars(fn))
    ns.update(_moog_any=_sym_any, _moog_all=_sym_all, _moog_not=_sym_not, _moog_and=_sym_and, _moog_or=_sym_or)
    if isinstance(node, ast.Lambda):
        expr = ast.Expression(node)
        ast.fix_missing_locations(expr)
        return eval(compile(expr, '<lowered>', 'eval'), ns)  # pylint: disable=eval-used
    node.decorator_list = []
    mod = ast.Module(body=[node], type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, '<lowered>', 'exec'), ns)  # pylint: disable=exec-used
    return ns[node.name]


_NP_RANDOM_FNS = ('uniform', 'rand', 'randn', 'randint', 'random', 'random_sample', 'ranf', 'sample', 'choice',
                  'normal', 'standard_normal', 'permutation', 'shuffle', 'binomial', 'poisson', 'exponential',
                  'beta', 'gamma', 'triangular', 'vonmises', 'laplace', 'lognormal', 'bytes')
_PY_RANDOM_FNS = ('random', 'uniform', 'randint', 'randrange', 'choice', 'choices', 'sample', 'shuffle', 'gauss',
                  'normalvariate', 'triangular', 'betavariate', 'expovariate', 'getrandbits')


class ImpureCallable(LoweringError):
    """A config callable drew a random number while it was being traced."""


@contextlib.contextmanager
def no_randomness(what='callable'):
    """While a config callable is traced ONCE into a device expression, any random draw it makes
    would be frozen into a constant for every env, step and episode (e.g. match_to_sample.py:227-228
    `s.velocity = np.random.uniform(-0.25, 0.25, size=(2,))`).  Inside this context the module-level
    draws of `numpy.random` and `random` raise instead, so such a config is refused loudly."""
    import random as _pyrandom

    def _refuse(name):
        def _raise(*args, **kwargs):
            raise ImpureCallable(
                '{} calls {} while being lowered: a random draw cannot be traced to a device '
                'expression (it would be frozen to one value for every env and step)'.format(what, name))
        return _raise

    saved = []
    for mod, names, prefix in ((np.random, _NP_RANDOM_FNS, 'numpy.random.'), (_pyrandom, _PY_RANDOM_FNS, 'random.')):
        for n in names:
            if hasattr(mod, n):
                saved.append((mod, n, getattr(mod, n)))
                setattr(mod, n, _refuse(prefix + n))
    try:
        yield
    finally:
        for mod, n, fn in saved:
            setattr(mod, n, fn)


def _symbolic_call(fn, *args, **options):
    """fn(*args) on symbolic sprites; when the callable uses Python's `and` / `or` / `not` /
    `any` / `all` on per-sprite values (which plain tracing cannot see), it is recompiled with
    those constructs turned into expression builders and traced again; when it still branches
    (`if` / `elif` / conditional expressions on a traced value) and has no effects
    (`branches=True`: predicates, conditions, rewards), every path is traced and the results are
    folded into selects.  Random draws are refused (`no_randomness`)."""
    branches = options.get('branches', True)
    with no_randomness(repr(getattr(fn, '__name__', fn))):
        try:
            return fn(*args)
        except ImpureCallable:
            raise
        except (LoweringError, TypeError):
            try:
                rewritten = _rewritten(fn)
            except LoweringError:
                raise
            except Exception as exc:  # pylint: disable=broad-except
                raise LoweringError('cannot lower {!r} to a device expression ({}: {})'.format(
                    getattr(fn, '__name__', fn), type(exc).__name__, exc))
            try:
                try:
                    return rewritten(*args)
                except LoweringError:
                    if not branches:
                        raise
                    return _explore(rewritten, args)
            except LoweringError:
                raise
            except Exception as exc:  # pylint: disable=broad-except
                raise LoweringError('cannot lower {!r} to a device expression ({}: {})'.format(
                    getattr(fn, '__name__', fn), type(exc).__name__, exc))


def compile_sprite_predicate(fn):
    """`fn(sprite) -> bool`  ->  postfix code, or None for "always true"."""
    if fn is None:
        return None
    out = _symbolic_call(fn, SymSprite(0))
    if out is True:
        return None
    return _code_of(out)


def compile_pair_condition(fn):
    """`fn(s0, s1[, meta_state]) -> bool` -> postfix code / None."""
    fn = _unwrap(fn, 'condition')
    if fn is None:
        return None
    n = _n_params(fn)
    args = (SymSprite(0), SymSprite(1)) + ((None,) if n >= 3 else ())
    out = _symbolic_call(fn, *args)
    if out is True:
        return None
    return _code_of(out)


def compile_modifier(fn):
    """`fn(sprite)` mutating the sprite -> postfix code of its stores."""
    stores = []
    _symbolic_call(fn, SymSprite(0, stores), branches=False)
    return _stores_to_code(stores)


def _stores_to_code(stores):
    """[(attribute name, value)] recorded while tracing -> postfix code of the stores (+ a final constant)."""
    code = []
    for name, value in stores:
        if name == 'position':
            code += value[0].code + value[1].code + [(X_STORE_POS, 0, 0.0)]
            continue
        if name == 'velocity':
            # Python evaluates the whole right-hand side first (`s.velocity = np.array([-s.y_vel, s.x_vel])`):
            # both components are pushed, then stored from the top of the stack down.  The setter installs
            # a FRESH array (sprite.py:639-643): one that is built from constants / positions only is
            # float64 and shared with nobody (c = 3: the device drops MOOG_SF_VEL32 and the alias id);
            # one computed from the sprite's own velocity keeps that velocity's dtype (c = 0).
            dtype_free = not any(op in (X_ATTR0, X_ATTR1) and (arg & 0xff) < len(ATTRS) and ATTRS[arg & 0xff] in ('x_vel', 'y_vel', 'angle_vel', 'angle')
                                 for v in value for op, arg, _ in v.code)
            kind = 3.0 if dtype_free else 0.0
            code += value[0].code + value[1].code + [(X_STORE, ATTRS.index('y_vel'), kind),
                                                     (X_STORE, ATTRS.index('x_vel'), kind)]
            continue
        # c: NumPy kind of the stored value as far as the device tracks it (angle: a pure constant is a
        # python float, anything computed from sprite factors counts as np.float64)
        reads = {ATTRS[arg & 0xff] if (arg & 0xff) < len(ATTRS) else 'metadata'
                 for op, arg, _ in value.code if op in (X_ATTR0, X_ATTR1)}
        weak = {'scale', 'aspect_ratio', 'mass', 'c0', 'c1', 'c2', 'opacity'}     # python numbers in the reference's Sprite
        if name == 'angle' and reads and reads <= weak | {'angle'} and 'angle' in reads:
            kind = 4.0      # `s.angle = s.angle + 0.5 * np.pi`: stays what the angle is (a python float at birth)
        else:
            kind = 2.0 if reads else 0.0
        code += value.code + [(X_STORE, ATTRS.index(name), kind)]
    # Leave a value on the stack so the VM has a defined result.
    return code + [(X_CONST, 0, 1.0)]


def compile_dependent_fn(fn, indep_keys, dep_keys):
    """`DependentDistribution.dependent_fn(sample_dict) -> dict` traced over the independent factors:
    {dependent key: (postfix code reading factors as X_ATTR0, set of attribute indices it reads)}."""
    sample = {}
    for k in indep_keys:
        if k not in ATTRS:
            raise LoweringError('DependentDistribution over the factor {!r} is not on the device sampler'.format(k))
        sample[k] = Sym([(X_ATTR0, ATTRS.index(k), 0.0)])
    with no_randomness('DependentDistribution dependent_fn'):
        out = fn(sample)
    codes = {}
    for k in dep_keys:
        if k not in ATTRS:
            raise LoweringError('a dependent factor {!r} is not on the device sampler'.format(k))
        v = Sym.lift(out[k])
        codes[k] = (v.code, {arg for op, arg, _ in v.code if op == X_ATTR0})
    return codes


def constant_reward(reward_fn):
    """ContactReward reward_fn -> float (contact_reward.py:48-51)."""
    if not callable(reward_fn):
        return float(reward_fn)
    cv = closure_vars(reward_fn)
    if set(cv) == {'reward_fn'} and not callable(cv['reward_fn']):
        return float(cv['reward_fn'])
    try:
        with no_randomness('ContactReward reward_fn'):
            out = reward_fn(SymSprite(0), SymSprite(1))
        return float(out)
    except ImpureCallable:
        raise
    except (LoweringError, TypeError):
        raise LoweringError(
            'ContactReward reward_fn must be a constant on the device path')


def pair_reward(reward_fn):
    """ContactReward reward_fn -> (constant, None) or (0.0, postfix code of
    `reward_fn(sprite_0, sprite_1)` over the two sprites' factors)."""
    try:
        return constant_reward(reward_fn), None
    except LoweringError:
        pass
    fn = _unwrap(reward_fn, 'reward_fn')
    try:
        out = _symbolic_call(fn, SymSprite(0), SymSprite(1))
    except LoweringError:
        raise
    except Exception as exc:  # pylint: disable=broad-except
        raise LoweringError(
            'ContactReward reward_fn cannot be traced to an expression over the two sprites\' '
            'factors ({}: {})'.format(type(exc).__name__, exc))
    if not isinstance(out, Sym):
        raise LoweringError('ContactReward reward_fn must return a number or an expression of sprite factors')
    return 0.0, out.code


def constant_state_reward(reward_fn):
    """Reset reward_fn -> float (reset.py:41-43 defaults to `lambda _: 0.`)."""
    if reward_fn is None:
        return 0.0
    try:
        with no_randomness('Reset reward_fn'):
            return float(reward_fn(None))
    except ImpureCallable:
        raise
    except Exception:  # pylint: disable=broad-except
        raise LoweringError(
            'Reset reward_fn must not depend on the state on the device path')


# ---------------------------------------------------------------------------
# state-level conditions
# ---------------------------------------------------------------------------

def contact_layers(fn):
    """(layer_0, layer_1) of a get_contact_indices / get_contact_counter."""
    if hasattr(fn, 'layer_0') and hasattr(fn, 'layer_1'):
        return fn.layer_0, fn.layer_1
    cv = closure_vars(fn)
    if {'layer_0', 'layer_1'} <= set(cv):
        return cv['layer_0'], cv['layer_1']
    if '_get_contact_indices' in cv:
        return contact_layers(cv['_get_contact_indices'])
    return None


def _lambda_ast(fn):
    """Parse the source of `fn` and return its Lambda / FunctionDef node."""
    try:
        src = textwrap.dedent(inspect.getsource(fn))
    except (OSError, TypeError):
        raise LoweringError('no source available for {!r}'.format(fn))
    if fn.__name__ != '<lambda>':
        tree = ast.parse(src)
        for node in ast.walk(tree):
            if isinstance(node, ast.FunctionDef) and node.name == fn.__name__:
                return node
        raise LoweringError('cannot find def {} in its source'.format(
            fn.__name__))
    # A lambda's "source" is the physical line(s) it sits on: cut at each
    # 'lambda' keyword and shrink from the right until it parses.
    argnames = list(inspect.signature(fn).parameters)
    pos = -1
    while True:
        pos = src.find('lambda', pos + 1)
        if pos < 0:
            break
        for end in range(len(src), pos + 6, -1):
            try:
                node = ast.parse(src[pos:end].strip(), mode='eval').body
            except SyntaxError:
                continue
            if isinstance(node, ast.Lambda) and [
                    a.arg for a in node.args.args] == argnames:
                return node
    raise LoweringError('cannot isolate the lambda in: {}'.format(src.strip()))


class _StateLowering(object):
    def __init__(self, fn, prog, node):
        self.fn = fn
        self.prog = prog
        self.ns = dict(getattr(fn, '__globals__', {}))
        self.ns.update(closure_vars(fn))
        args = [a.arg for a in node.args.args]
        self.state_name = args[0]

    def fail(self, node):
        raise LoweringError(
            'state condition is not in the supported form (all / any / sum / '
            'len over one layer, combined with not / and / or / comparisons): '
            + ast.unparse(node))

    def _layer_of(self, node):
        """state['name'] -> 'name'."""
        if (isinstance(node, ast.Subscript) and isinstance(node.value, ast.Name)
                and node.value.id == self.state_name):
            key = node.slice
            try:
                return eval(compile(ast.Expression(key), '<cond>', 'eval'),  # pylint: disable=eval-used
                            self.ns)
            except Exception:  # pylint: disable=broad-except
                self.fail(node)
        self.fail(node)

    def _comprehension(self, node):
        """[elt for s in state['L'] (if cond)] -> (layer, Sym)."""
        if not isinstance(node, (ast.ListComp, ast.GeneratorExp)):
            self.fail(node)
        if len(node.generators) != 1:
            self.fail(node)
        gen = node.generators[0]
        if not isinstance(gen.target, ast.Name):
            self.fail(node)
        layer = self._layer_of(gen.iter)
        var = gen.target.id

        def _trace(expr_node):
            fn_node = ast.Expression(ast.Lambda(
                args=ast.arguments(posonlyargs=[], args=[ast.arg(arg=var)],
                                   kwonlyargs=[], kw_defaults=[], defaults=[]),
                body=expr_node))
            ast.fix_missing_locations(fn_node)
            f = eval(compile(fn_node, '<cond>', 'eval'), self.ns)  # pylint: disable=eval-used
            return Sym.lift(f(SymSprite(0)))

        elt = _trace(node.elt)
        for cond in gen.ifs:
            elt = elt & _trace(cond)
        return layer, elt

    def lower(self, node):
        """-> op index producing the value of `node` as a double."""
        prog = self.prog
        # the C enum values are duplicated here to avoid a circular import
        SC_ALL, SC_ANY, SC_COUNT, SC_CONST = 160, 161, 162, 165
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name):
            name = node.func.id
            if name in ('all', 'any', 'sum') and len(node.args) == 1:
                layer, elt = self._comprehension(node.args[0])
                ls, ln = prog.add_list([layer])
                kind = {'all': SC_ALL, 'any': SC_ANY, 'sum': SC_COUNT}[name]
                return ('op', prog.emit(kind, 0, (ls, ln, prog.add_expr(elt.code))))
            if name == 'len' and len(node.args) == 1:
                layer = self._layer_of(node.args[0])
                ls, ln = prog.add_list([layer])
                return ('op', prog.emit(SC_COUNT, 0, (ls, ln, -1)))
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name):
            known = self._known_helper(node)
            if known is not None:
                return known
        if isinstance(node, ast.Constant) and isinstance(
                node.value, (bool, int, float)):
            return ('const', float(node.value))
        # comparisons / not / and / or / + - * over the forms above
        SC_BINARY, SC_NOT = 166, 167
        if isinstance(node, ast.Compare) and len(node.ops) == 1:
            x = {ast.Lt: X_LT, ast.LtE: X_LE, ast.Gt: X_GT, ast.GtE: X_GE,
                 ast.Eq: X_EQ, ast.NotEq: X_NE}.get(type(node.ops[0]))
            if x is not None:
                a = self._as_op(self.lower(node.left))
                b = self._as_op(self.lower(node.comparators[0]))
                return ('op', prog.emit(SC_BINARY, 0, (a, b, x)))
        if isinstance(node, ast.BoolOp):
            x = X_AND if isinstance(node.op, ast.And) else X_OR
            acc = self._as_op(self.lower(node.values[0]))
            for v in node.values[1:]:
                acc = prog.emit(SC_BINARY, 0,
                                (acc, self._as_op(self.lower(v)), x))
            return ('op', acc)
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, ast.Not):
            return ('op', prog.emit(
                SC_NOT, 0, (self._as_op(self.lower(node.operand)),)))
        if isinstance(node, ast.BinOp):
            x = {ast.Add: X_ADD, ast.Sub: X_SUB, ast.Mult: X_MUL}.get(
                type(node.op))
            if x is not None:
                a = self._as_op(self.lower(node.left))
                b = self._as_op(self.lower(node.right))
                return ('op', prog.emit(SC_BINARY, 0, (a, b, x)))
        if isinstance(node, ast.Name) and node.id in self.ns and isinstance(
                self.ns[node.id], (bool, int, float)):
            return ('const', float(self.ns[node.id]))
        bern = self._bernoulli(node)
        if bern is not None:
            return bern
        first = self._trace_first(node)
        if first is not None:
            return first
        self.fail(node)

    def _bernoulli(self, node):
        """`np.random.binomial(1, p)` with a constant p (first_person_predators_prey.py:181,190) ->
        MOOG_SC_BERNOULLI: true when the step's uniform of a rule-noise column of its own is < p."""
        if not isinstance(node, ast.Call):
            return None
        try:
            func = eval(compile(ast.Expression(node.func), '<cond>', 'eval'), self.ns)  # pylint: disable=eval-used
        except Exception:  # pylint: disable=broad-except
            return None
        if func is not np.random.binomial:
            return None
        try:
            args = [eval(compile(ast.Expression(a), '<cond>', 'eval'), self.ns) for a in node.args]  # pylint: disable=eval-used
            kwargs = {k.arg: eval(compile(ast.Expression(k.value), '<cond>', 'eval'), self.ns)  # pylint: disable=eval-used
                      for k in node.keywords}
        except Exception:  # pylint: disable=broad-except
            self.fail(node)
        names = ('n', 'p', 'size')
        for k, v in zip(names, args):
            kwargs[k] = v
        if set(kwargs) - set(names) or kwargs.get('size') is not None or kwargs.get('n') != 1 or 'p' not in kwargs:
            self.fail(node)
        p = float(kwargs['p'])
        if not 0.0 <= p <= 1.0:
            self.fail(node)
        col = self.prog.rule_noise_dim
        self.prog.rule_noise_dim += 1
        return ('op', self.prog.emit(169, 0, (col,), (p,)))  # MOOG_SC_BERNOULLI

    def _trace_first(self, node):
        """An expression over `state[layer][0]` (one layer), e.g. pacman.py's
        `np.all(state['agent'][0].velocity == 0)` -> MOOG_SC_FIRST."""
        sym_state = SymState()
        ns = dict(self.ns)
        ns[self.state_name] = sym_state
        try:
            expr = ast.Expression(node)
            ast.fix_missing_locations(expr)
            out = eval(compile(expr, '<cond>', 'eval'), ns)  # pylint: disable=eval-used
        except LoweringError:
            raise
        except Exception:  # pylint: disable=broad-except
            return None
        if isinstance(out, SymVec):
            return None
        layers = set(sym_state.layers)
        if len(layers) != 1 or not isinstance(out, Sym):
            return None
        ls, ln = self.prog.add_list([layers.pop()])
        return ('op', self.prog.emit(168, 0, (ls, ln, self.prog.add_expr(out.code))))  # MOOG_SC_FIRST

    def _as_op(self, lowered):
        kind, value = lowered
        if kind == 'const':
            return self.prog.emit(165, 0, (), (value,))  # MOOG_SC_CONST
        return value

    # -- helpers of shipped configs that cannot be traced (data-dependent
    #    Python control flow) and have a declarative device equivalent ---------
    _AGENTS_CONTACTING_LAYER = (
        "def agents_contacting_layer(state, layer, value):\n"
        "    n_contact = 0\n"
        "    for s in state[layer]:\n"
        "        if s.c2 != value:\n"
        "            continue\n"
        "        n_contact += s.overlaps_sprite(state['agent_0'][0]) or "
        "s.overlaps_sprite(state['agent_1'][0]) or "
        "s.overlaps_sprite(state['agent_2'][0])\n"
        "    return n_contact")

    def _known_helper(self, node):
        """`agents_contacting_layer(state, layer, value)` of cleanup.py:183-193:
        the number of `layer` sprites with c2 == value that overlap any agent
        (Python `or`: the agents are tried in order until one overlaps) ->
        MOOG_SC_CONTACT_ANY_COUNT.  The helper's source must match verbatim
        (modulo formatting), otherwise the condition is rejected."""
        fn = self.ns.get(node.func.id)
        if fn is None or getattr(fn, '__name__', '') != 'agents_contacting_layer':
            return None
        try:
            src = textwrap.dedent(inspect.getsource(fn))
            got = ast.dump(ast.parse(src).body[0])
            want = ast.dump(ast.parse(self._AGENTS_CONTACTING_LAYER).body[0])
        except (OSError, TypeError, SyntaxError):
            return None
        if got != want or len(node.args) != 3:
            return None
        if not (isinstance(node.args[0], ast.Name) and node.args[0].id == self.state_name):
            return None
        try:
            layer, value = [eval(compile(ast.Expression(a), '<cond>', 'eval'), self.ns)  # pylint: disable=eval-used
                            for a in node.args[1:]]
        except Exception:  # pylint: disable=broad-except
            return None
        SC_CONTACT_ANY_COUNT = 164
        prog = self.prog
        ls, ln = prog.add_list([layer])
        as_, an = prog.add_list(['agent_0', 'agent_1', 'agent_2'])
        filt = (SymSprite(0).c2 == float(value))
        return ('op', prog.emit(SC_CONTACT_ANY_COUNT, 0,
                                (ls, ln, as_, an, prog.add_expr(Sym.lift(filt).code))))


def compile_state_condition(cond, prog):
    """State condition -> index of the op that evaluates it.

    Supported: contact counters (`get_contact_counter(l0, l1)`); and Python
    callables of the form `all(...)`, `any(...)`, `sum(...)`, `len(state[L])`
    over one layer with a per-sprite expression.
    """
    from . import compiler as C  # late import: compiler imports this module
    cond = _unwrap(cond, 'condition')
    layers = contact_layers(cond) if callable(cond) else None
    if layers is not None:
        return prog.emit(C.SC_CONTACT_COUNT, 0,
                         (prog.layer_index(layers[0]),
                          prog.layer_index(layers[1])))
    declared = getattr(cond, 'moog_b200_condition', None)
    if declared is not None:
        return declared(prog)
    try:
        takes_meta = _n_params(cond) >= 2
    except (TypeError, ValueError):
        takes_meta = False
    if takes_meta:      # (state, meta_state): the env's meta_state entries are variables of the record
        return state_tree(cond, prog, 'state condition {}'.format(getattr(cond, '__name__', cond)))
    node = _lambda_ast(cond)
    body = node.body
    if isinstance(body, list):  # def: single `return <expr>`
        stmts = [s for s in body if not (
            isinstance(s, ast.Expr) and isinstance(s.value, ast.Constant))]
        if len(stmts) != 1 or not isinstance(stmts[0], ast.Return):
            raise LoweringError(
                'state condition {} must be a single return expression'.format(
                    cond.__name__))
        body = stmts[0].value
    low = _StateLowering(cond, prog, node)
    n_ops, n_expr, n_ipool = len(prog.ops), len(prog.expr), len(prog.ipool)
    try:
        kind, value = low.lower(body)
    except LoweringError as first:
        # not one of the aggregate forms: a callable that picks single sprites and branches on them
        del prog.ops[n_ops:], prog.expr[n_expr:], prog.ipool[n_ipool:]
        try:
            return state_tree(cond, prog, 'state condition {}'.format(getattr(cond, '__name__', cond)))
        except LoweringError as second:
            raise LoweringError('{}; as a decision tree: {}'.format(first, second))
    if kind == 'const':
        return prog.emit(C.SC_CONST, 0, (), (value,))
    return value


def state_reward(reward_fn, prog):
    """Reset reward_fn (reset.py:41-43, called only when the condition holds) -> (constant, None), or
    (0.0, index of the MOOG_SC_TREE op that evaluates it)."""
    try:
        return constant_state_reward(reward_fn), None
    except ImpureCallable:
        raise
    except LoweringError:
        return 0.0, state_tree(reward_fn, prog, 'Reset reward_fn')
