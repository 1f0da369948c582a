"""CPU oracle of anchor decode + objectness threshold + per-class NMS.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates utility/utils.py:208-257 ``decode_netout`` with its helpers ``sigmoid`` :259-260,
``softmax`` :262-270 (GLOBAL max subtraction and the global-min < -100 temperature rescale),
``bbox_iou`` :155-173 and ``interval_overlap`` :175-188.  Arithmetic type follows what the
reference does under numpy >= 2 on a float32 ``netout``: every tensor op and every per-box
scalar op is float32 (python floats are weak scalars under NEP 50).

Pinned bit-exactly against the reference's own function exec'd from /root/reference by
oracle/make_golden.py (thousands of seeded + adversarial tensors); the committed
tests/golden/decode_*.npz carry the reference's outputs.

Tie rule: the reference orders each class by ``reversed(np.argsort(p))`` whose order among
equal values is unspecified (quicksort).  Equal *zero* entries are no-ops in the greedy loop, so
only ties among equal non-zero probabilities are order-dependent; this oracle (and the CUDA
kernel) resolve them as a stable ascending sort reversed, i.e. larger candidate index first.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


class Box:
    """Mirror of utility/utils.py:113-136 ``BoundBox`` (centre-form, image-relative)."""
    __slots__ = ("x", "y", "w", "h", "c", "classes", "cell", "_label", "_score")

    def __init__(self, x, y, w, h, c, classes, cell=-1):
        self.x, self.y, self.w, self.h, self.c = x, y, w, h, c
        self.classes = classes
        self.cell = cell                    # (row*grid_w + col)*nb_box + b : the candidate's anchor id
        self._label = -1
        self._score = -1

    def get_label(self) -> int:
        if self._label == -1:
            self._label = int(np.argmax(self.classes))
        return self._label

    def get_score(self):
        if self._score == -1:
            self._score = self.classes[self.get_label()]
        return self._score


def _sigmoid(x):
    return 1. / (1. + np.exp(-x))


def class_probabilities(netout: np.ndarray, obj_threshold: float) -> np.ndarray:
    """Steps (1)-(3): returns (conf, p) with p = conf*softmax(cls) gated by ``> obj_threshold``."""
    conf = _sigmoid(netout[..., 4])
    z = netout[..., 5:]
    z = z - np.max(z)                                   # GLOBAL max over the whole tensor
    lo = np.min(z)
    if lo < -100.:
        z = z / lo * -100.                              # global temperature rescale
    e = np.exp(z)
    p = conf[..., np.newaxis] * (e / e.sum(-1, keepdims=True))
    p = p * (p > obj_threshold)
    return conf, p


def _overlap(a_lo, a_hi, b_lo, b_hi):
    if b_lo < a_lo:
        return 0 if b_hi < a_lo else min(a_hi, b_hi) - a_lo
    return 0 if a_hi < b_lo else min(a_hi, b_hi) - b_lo


def box_iou(p: Box, q: Box):
    iw = _overlap(p.x - p.w / 2, p.x + p.w / 2, q.x - q.w / 2, q.x + q.w / 2)
    ih = _overlap(p.y - p.h / 2, p.y + p.h / 2, q.y - q.h / 2, q.y + q.h / 2)
    inter = iw * ih
    union = p.w * p.h + q.w * q.h - inter
    return float(inter) / union


def decode_netout(netout: np.ndarray, obj_threshold: float, nms_threshold: float,
                  anchors: Sequence[float], nb_class: int) -> List[Box]:
    """netout: (grid_h, grid_w, nb_box, 5+nb_class) float32 raw conv_23 logits (not mutated)."""
    net = np.array(netout, dtype=np.float32, copy=True)
    grid_h, grid_w, nb_box = net.shape[:3]
    conf, prob = class_probabilities(net, obj_threshold)

    boxes: List[Box] = []
    for row, col, b in np.argwhere(prob.any(axis=-1)):           # row-major (row, col, b) order
        row, col, b = int(row), int(col), int(b)
        tx, ty, tw, th = net[row, col, b, :4]
        x = (col + _sigmoid(tx)) / grid_w
        y = (row + _sigmoid(ty)) / grid_h
        w = anchors[2 * b + 0] * np.exp(tw) / grid_w
        h = anchors[2 * b + 1] * np.exp(th) / grid_h
        boxes.append(Box(x, y, w, h, conf[row, col, b], prob[row, col, b].copy(),
                         cell=(row * grid_w + col) * nb_box + b))

    n = len(boxes)
    for c in range(nb_class):
        col_p = np.array([bx.classes[c] for bx in boxes], dtype=np.float32)
        live = np.nonzero(col_p)[0]
        if live.size == 0:
            continue
        # value descending; equal values: larger candidate index first (reversed stable argsort)
        order = sorted(live.tolist(), key=lambda i: (-float(col_p[i]), -i))
        for a in range(len(order)):
            i = order[a]
            if boxes[i].classes[c] == 0:
                continue
            for bidx in range(a + 1, len(order)):
                j = order[bidx]
                if box_iou(boxes[i], boxes[j]) >= nms_threshold:
                    boxes[j].classes[c] = 0
    assert n == len(boxes)
    return [bx for bx in boxes if bx.get_score() > obj_threshold]


def boxes_to_array(boxes: List[Box]) -> np.ndarray:
    """(n, 8) float64 rows [x, y, w, h, conf, score, label, cell] -- the record the C-ABI emits."""
    out = np.zeros((len(boxes), 8), np.float64)
    for i, b in enumerate(boxes):
        out[i] = (b.x, b.y, b.w, b.h, b.c, b.get_score(), b.get_label(), b.cell)
    return out
