"""Colour maps named by PILRenderer(color_to_rgb=...) (color_maps.py:21-23).

The device rasteriser implements the same map in-kernel (float64 colorsys
formula, x255, truncation to uint8); this host copy serves single colours.
"""

import colorsys

import numpy as np


def hsv_to_rgb(c):
    r, g, b = colorsys.hsv_to_rgb(*c)
    return tuple(np.array([255 * r, 255 * g, 255 * b]).astype(np.uint8))
