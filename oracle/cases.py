"""Seeded input generators shared by oracle/make_golden.py and tests/.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import numpy as np

DECODE_SPECS = [(13, 80), (13, 20), (13, 12), (13, 2), (19, 80)]
DECODE_KINDS = ["empty", "few", "crowd", "cluster", "rescale", "plain"]


def decode_case(seed, g, c, kind):
    """Random conv_23-like logits with a controllable number of confident anchors."""
    rng = np.random.default_rng(seed)
    net = rng.standard_normal((g, g, 5, 5 + c)).astype(np.float32)
    net[..., 4] -= 3.0
    if kind == "empty":
        net[..., 4] -= 10.0
    elif kind in ("few", "crowd", "cluster", "rescale"):
        n_obj = {"few": 6, "crowd": 120, "cluster": 40, "rescale": 12}[kind]
        for _ in range(n_obj):
            if kind == "cluster":
                r, q = 6 + int(rng.integers(-1, 2)), 6 + int(rng.integers(-1, 2))
            else:
                r, q = int(rng.integers(0, g)), int(rng.integers(0, g))
            b = int(rng.integers(0, 5))
            net[r, q, b, 4] = rng.uniform(2.0, 6.0)
            k = int(rng.integers(0, min(c, 3)))
            net[r, q, b, 5 + k] += rng.uniform(6.0, 10.0)
            net[r, q, b, 2:4] = rng.uniform(-0.5, 0.8, 2)
        if kind == "rescale":
            net[0, 0, 0, 5] = -150.0            # global min < -100 after max-subtraction
    return net


def heatmap_case_inputs(n: int, seed: int):
    """Seeded (x, y, w, h) rows as preprocessing.py:452-456 forms them (top-left corner, may be negative or past
    the frame) and thresholded-noise heat-maps for the read-back."""
    rng = np.random.default_rng(seed)
    xywh = rng.uniform(-0.3, 1.2, (n, 4))
    xywh[:, 2:] = np.abs(rng.uniform(0, 0.9, (n, 2)))
    xywh[0] = 0.0                                       # the "no detection" row (preprocessing.py:445-449)
    xywh[1] = (0.5, 0.5, 0.0, 0.0)
    xywh[2] = (0.999, 0.999, 0.5, 0.5)
    xywh[3] = (-0.01, -0.01, 0.1, 0.1)                  # int() truncates toward zero: -0.32 -> 0
    xywh[4] = (-0.5, 0.2, 0.3, 0.3)                     # negative slice start wraps once in numpy
    heat = rng.uniform(0, 1, (n, 32, 32)) * (rng.uniform(0, 1, (n, 1, 1)) ** 2 + 0.6)
    heat[0] = 0.0                                       # nothing >= thresh -> (32, 32, -1, -1)
    heat[1] = 1.0
    # float32-representable values: the device entry points take float32 rows and the python int() truncation must
    # see the same numbers
    return xywh.astype(np.float32).astype(np.float64), heat.astype(np.float32).astype(np.float64)
