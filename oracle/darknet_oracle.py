"""CPU oracle of the darknet region layer + box decode + objectness NMS (the path behind
models_detection/YOLO.py:140-162).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates darknet/src/region_layer.c:158-185 (forward_region_layer, CPU branch),
region_layer.c:76-84 (get_region_box), :364-437 (get_region_detections, softmax=1, no tree,
no map), :336-362 (correct_region_boxes), darknet/src/box.c:21-55 (do_nms_obj), :152-182
(box_iou) and blas.c softmax_cpu.  Pinned against oracle/_ref/libdarknet.so by make_golden.py.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

ANCHORS = [0.57273, 0.677385, 1.87446, 2.06253, 3.33843, 5.47434, 7.88282, 3.52778, 9.77052, 9.16828]


def _logistic(x):
    return (1. / (1. + np.exp(-x.astype(np.float64)))).astype(np.float32)


def region_forward(logits_chw: np.ndarray, n_class: int, n_box: int = 5) -> np.ndarray:
    """(A*(5+C), G, G) raw conv output -> region layer output, same layout
    (entry_index: channel = a*(5+C) + entry)."""
    d = 5 + n_class
    g_h, g_w = logits_chw.shape[1:]
    x = np.array(logits_chw, dtype=np.float32).reshape(n_box, d, g_h, g_w)
    x[:, 0:2] = _logistic(x[:, 0:2])
    x[:, 4] = _logistic(x[:, 4])
    cls = x[:, 5:].astype(np.float64)
    e = np.exp(cls - cls.max(axis=1, keepdims=True))
    x[:, 5:] = (e / e.sum(axis=1, keepdims=True)).astype(np.float32)
    return x.reshape(n_box * d, g_h, g_w)


def _iou(a, b) -> float:
    def ov(x1, w1, x2, w2):
        return min(x1 + w1 / 2, x2 + w2 / 2) - max(x1 - w1 / 2, x2 - w2 / 2)
    w = ov(a[0], a[2], b[0], b[2])
    h = ov(a[1], a[3], b[1], b[3])
    inter = 0.0 if (w < 0 or h < 0) else w * h
    return np.float32(inter) / np.float32(a[2] * a[3] + b[2] * b[3] - inter)


def detect(region_chw: np.ndarray, im_w: int, im_h: int, net_w: int, net_h: int, thresh: float,
           nms: float, n_class: int, n_box: int = 5) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """-> (boxes (n,4) [cx,cy,w,h] in image pixels, objectness (n,), prob (n,C)), n = G*G*A, in
    the library's pre-sort index order  index = a*G*G + row*G + col ; suppressed rows zeroed."""
    d = 5 + n_class
    g_h, g_w = region_chw.shape[1:]
    r = region_chw.reshape(n_box, d, g_h, g_w)
    n = n_box * g_h * g_w
    boxes = np.zeros((n, 4), np.float32)
    obj = np.zeros(n, np.float32)
    prob = np.zeros((n, n_class), np.float32)
    if (float(net_w) / im_w) < (float(net_h) / im_h):
        new_w, new_h = net_w, (im_h * net_w) // im_w
    else:
        new_h, new_w = net_h, (im_w * net_h) // im_h
    for a in range(n_box):
        for row in range(g_h):
            for col in range(g_w):
                i = a * g_h * g_w + row * g_w + col
                bx = np.float32((col + r[a, 0, row, col]) / np.float32(g_w))
                by = np.float32((row + r[a, 1, row, col]) / np.float32(g_h))
                bw = np.float32(np.exp(np.float64(r[a, 2, row, col])) * np.float32(ANCHORS[2 * a]) / g_w)
                bh = np.float32(np.exp(np.float64(r[a, 3, row, col])) * np.float32(ANCHORS[2 * a + 1]) / g_h)
                # correct_region_boxes, relative = 0
                bx = np.float32((bx - (net_w - new_w) / 2. / net_w) / np.float32(np.float32(new_w) / net_w))
                by = np.float32((by - (net_h - new_h) / 2. / net_h) / np.float32(np.float32(new_h) / net_h))
                bw = np.float32(bw * np.float32(np.float32(net_w) / new_w))
                bh = np.float32(bh * np.float32(np.float32(net_h) / new_h))
                boxes[i] = (bx * im_w, by * im_h, bw * im_w, bh * im_h)
                s = r[a, 4, row, col]
                if s > np.float32(thresh):
                    obj[i] = s
                    p = s * r[a, 5:, row, col]
                    prob[i] = np.where(p > np.float32(thresh), p, 0)
    if nms:
        order = [i for i in np.argsort(-obj, kind="stable") if obj[i] != 0]
        for ai, i in enumerate(order):
            if obj[i] == 0:
                continue
            for j in order[ai + 1:]:
                if obj[j] == 0:
                    continue
                if _iou(boxes[i], boxes[j]) > np.float32(nms):
                    obj[j] = 0
                    prob[j] = 0
        order = [i for i in order if obj[i] != 0]
        rest = [i for i in range(n) if obj[i] == 0]
        perm = np.array(order + rest, dtype=np.int64)
        boxes, obj, prob = boxes[perm], obj[perm], prob[perm]
    return boxes, obj, prob


def yolo_detect_list(boxes, obj, prob, names):
    """models_detection/YOLO.py:152-159: [(name, prob, (x,y,w,h))] sorted by -prob."""
    res = []
    for j in range(len(obj)):
        for i in range(prob.shape[1]):
            if prob[j, i] > 0:
                res.append((names[i], float(prob[j, i]), tuple(float(v) for v in boxes[j])))
    return sorted(res, key=lambda t: -t[1])
