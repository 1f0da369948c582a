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


