"""Game-rule specs (reference: moog/game_rules/).

On the hot path (lowered to device ops by moog_b200.compiler):
`VanishOnContact`, `VanishByFilter`, `ModifyOnContact`, `ModifySprites`,
`ConditionalRule`, plus the `get_contact_indices` / `get_contact_counter`
condition builders, `TimedRule` / `DelayedRule` / `TemporaryRule` with fixed
intervals, `KeepNearCenter`, `Portal`, `ChangeLayer`, `CreateSprites`, `Fixation`, `Phase` / `PhaseSequence`,
`ModifyMetaState` / `UpdateMetaStateValue`.  A rule class of a config's own is traced by the compiler
(lambdas.trace_rule); what cannot be lowered raises at compile time, so a config never silently loses a rule.
"""

import abc

import numpy as np


def _as_tuple(x):
    return tuple(x) if isinstance(x, (list, tuple)) else (x,)


class AbstractRule(abc.ABC):
    def reset(self, state, meta_state):
        pass

    def step(self, state, meta_state):
        raise RuntimeError(
            'game rules are applied on the device by BatchedEnvironment')


class ContactCondition(object):
    """Declarative state -> int condition: number of contacting index pairs
    between two layers (contact_rules.py:15-51)."""

    def __init__(self, layer_0, layer_1, as_indices=False):
        self.layer_0 = layer_0
        self.layer_1 = layer_1
        self.as_indices = as_indices

    def __call__(self, state):
        raise RuntimeError('contact conditions are evaluated on the device')


def get_contact_indices(layer_0, layer_1):
    return ContactCondition(layer_0, layer_1, as_indices=True)


def get_contact_counter(layer_0, layer_1):
    return ContactCondition(layer_0, layer_1)


class ModifyOnContact(AbstractRule):
    """contact_rules.py:54-96."""

    def __init__(self, layers_0, layers_1, modifier_0=None, modifier_1=None,
                 filter_0=None, filter_1=None):
        self._layers_0 = _as_tuple(layers_0)
        self._layers_1 = _as_tuple(layers_1)
        self._modifier_0 = modifier_0
        self._modifier_1 = modifier_1
        self._filter_0 = filter_0
        self._filter_1 = filter_1


class Vanish(AbstractRule):
    def __init__(self, layer):
        self._layer = layer


class VanishByFilter(Vanish):
    """vanish.py:42-63."""

    def __init__(self, layer, filter_fn=None):
        super().__init__(layer)
        self._filter_fn = filter_fn


class VanishOnContact(Vanish):
    """vanish.py:66-86."""

    def __init__(self, vanishing_layer, contacting_layer):
        super().__init__(vanishing_layer)
        self._contacting_layer = contacting_layer
        self._get_contact_indices = get_contact_indices(
            vanishing_layer, contacting_layer)


class ModifySprites(AbstractRule):
    """modify_sprites.py:17-33."""

    def __init__(self, layers, modifier, sample_one=False, filter_fn=None):
        self._layers = [layers] if isinstance(layers, str) else layers
        self._modifier = modifier
        self._sample_one = sample_one
        self._filter_fn = filter_fn


class ConditionalRule(AbstractRule):
    """conditional.py:30-53."""

    def __init__(self, condition, rules):
        self._condition = condition
        self._rules = list(rules) if isinstance(rules, (list, tuple)) else [
            rules]


class TimedRule(AbstractRule):
    """Steps `rules` only while the call count since reset lies in
    `step_interval = (start, stop)` (timing.py:15-56)."""

    def __init__(self, step_interval, rules):
        self._step_interval = step_interval if callable(step_interval) else (lambda: step_interval)
        self._rules = list(rules) if isinstance(rules, (list, tuple)) else [rules]


class DelayedRule(TimedRule):
    """Starts after `steps_until_start` calls, runs for `duration` (timing.py:59-86)."""

    def __init__(self, steps_until_start, rules, duration=np.inf):
        start = steps_until_start if callable(steps_until_start) else (lambda: steps_until_start)
        length = duration if callable(duration) else (lambda: duration)

        def _interval():
            t0 = start()
            return (t0, t0 + length())
        super().__init__(_interval, rules)


class TemporaryRule(TimedRule):
    """Runs from the reset for `steps_until_stop` calls (timing.py:89-107)."""

    def __init__(self, steps_until_stop, rules):
        stop = steps_until_stop if callable(steps_until_stop) else (lambda: steps_until_stop)
        super().__init__(lambda: (0, stop()), rules)


class KeepNearCenter(AbstractRule):
    """Snaps the agent and the listed layers back by one grid cell whenever the agent is
    more than a cell away from (0.5, 0.5) (re_center.py:13-76)."""

    def __init__(self, agent_layer, layers_to_center, grid_x, grid_y=None):
        self._agent_layer = agent_layer
        layers_to_center = list(layers_to_center)
        if agent_layer not in set(layers_to_center):
            layers_to_center = layers_to_center + [agent_layer]
        self._layers_to_center = layers_to_center
        self._grid_cell = np.array([grid_x, grid_x if grid_y is None else grid_y])


class Portal(AbstractRule):
    """A sprite of `teleporting_layer` whose position enters a portal sprite reappears at the
    position of that portal's partner (portals are paired in order) and cannot teleport again
    until it has left every portal (portal.py:14-76)."""

    def __init__(self, teleporting_layer, portal_layer):
        self._teleporting_layer = teleporting_layer
        self._portal_layer = portal_layer


class ChangeLayer(AbstractRule):
    """Moves the sprites of `old_layer` that pass `filter_fn` to the end of `new_layer`
    (change_layer.py:11-45)."""

    def __init__(self, old_layer, new_layer, filter_fn=None):
        self._old_layer = old_layer
        self._new_layer = new_layer
        self._filter_fn = filter_fn if filter_fn is not None else (lambda s: True)


class CreateSprites(AbstractRule):
    """Appends the sprites `generator(without_overlapping=<the sprites of those layers>)` returns to
    `layer` (create_sprites.py:8-34).  `generator` must come from
    sprite_generators.generate_sprites: its factor distribution is what the device draws from."""

    def __init__(self, layer, generator, without_overlapping=()):
        self._layer = layer
        self._generator = generator
        self._without_overlapping = without_overlapping


class Fixation(AbstractRule):
    """Counts in `meta_state[meta_state_fixation_key]` for how many consecutive steps the first sprite of
    `agent_layer` has been within `fixation_threshold` of the first sprite of `fixation_layer`
    (fixation.py:19-54)."""

    def __init__(self, agent_layer, fixation_layer, fixation_threshold=0.1,
                 meta_state_fixation_key='fixation_duration'):
        self._agent_layer = agent_layer
        self._fixation_layer = fixation_layer
        self._fixation_threshold = fixation_threshold
        self._meta_state_fixation_key = meta_state_fixation_key


class Phase(AbstractRule):
    """`one_time_rules` on its first step, `continual_rules` on every step, until `duration` steps have
    passed or `end_condition(state[, meta_state])` holds (task_phases.py:18-95)."""

    def __init__(self, one_time_rules=(), continual_rules=(), end_condition=None, duration=np.inf, name=''):
        self._one_time_rules = one_time_rules if isinstance(one_time_rules, (list, tuple)) else (one_time_rules,)
        self._continual_rules = continual_rules if isinstance(continual_rules, (list, tuple)) else (continual_rules,)
        self._end_condition = end_condition if end_condition is not None else (lambda state, meta_state: False)
        self._duration = duration if callable(duration) else (lambda: duration)
        self._name = name

    @property
    def name(self):
        return self._name


class PhaseSequence(AbstractRule):
    """Phases stepped one after the other; `meta_state[meta_state_phase_name_key]` names the current one
    (task_phases.py:98-141)."""

    def __init__(self, *single_phases, meta_state_phase_name_key=None):
        self._phases = single_phases
        self._meta_state_key = meta_state_phase_name_key


class ModifyMetaState(AbstractRule):
    """`modifier(meta_state)` every step (modify_meta_state.py:8-26).  On the device the modifier is traced:
    it may read and assign numeric / string entries of the dict."""

    def __init__(self, modifier):
        self._modifier = modifier

    def step(self, state, meta_state):
        del state
        self._modifier(meta_state)


class UpdateMetaStateValue(AbstractRule):
    """`meta_state[key] = value` (modify_meta_state.py:29-48)."""

    def __init__(self, key, value):
        self._key = key
        self._value = value

    def step(self, state, meta_state):
        del state
        meta_state[self._key] = self._value
