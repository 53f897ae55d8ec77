"""Polygon-modifier specs (polygon_modifiers.py:32-98): how sprite outlines
are shifted / replicated before rasterisation.  Lowered to a render-kernel
mode by the config compiler."""

import abc


class AbstractPolygonModifier(abc.ABC):
    pass


class DoNothing(AbstractPolygonModifier):
    pass


class FirstPersonAgent(AbstractPolygonModifier):
    """Translate everything so the first sprite of `agent_layer` sits at
    (0.5, 0.5)."""

    def __init__(self, agent_layer):
        self._agent_layer = agent_layer


class TorusGeometry(AbstractPolygonModifier):
    """Draw every sprite at the 3x3 grid of unit offsets."""

    def __init__(self, wrap_layers):
        self._wrap_layers = (
            list(wrap_layers) if isinstance(wrap_layers, (list, tuple))
            else [wrap_layers])
