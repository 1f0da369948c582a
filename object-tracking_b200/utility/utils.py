"""Drop-in for the hot-path half of the reference's ``utility/utils.py`` (:113-270).

``decode_netout`` keeps the reference signature and return type (list of ``BoundBox``) but runs the
anchor decode + threshold + per-class NMS in one CUDA kernel (csrc/decode_nms.cu) through the C-ABI; the
small value types / scalar helpers (``BoundBox``, ``bbox_iou``, ``normalize``, ``sigmoid``, ``softmax``,
``WeightReader``) are plain host code with the reference's names and semantics so caller code keeps working.
Heat-map helpers (:53-79) run on the device as well.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np


class BoundBox:
    """utility/utils.py:113-136: centre-form, image-relative box with per-class scores."""

    def __init__(self, x, y, w, h, c=None, classes=None):
        self.x, self.y, self.w, self.h = x, y, w, h
        self.c = c
        self.classes = classes
        self.label = -1
        self.score = -1

    def get_label(self):
        if self.label == -1:
            self.label = int(np.argmax(self.classes))
        return self.label

    def get_score(self):
        if self.score == -1:
            self.score = self.classes[self.get_label()]
        return self.score


class WeightReader:
    """utility/utils.py:138-148: flat float32 reader that skips the 4-word header of a v0.1 file."""

    def __init__(self, weight_file):
        self.offset = 4
        self.all_weights = np.fromfile(weight_file, dtype="float32")

    def read_bytes(self, size):
        self.offset = self.offset + size
        return self.all_weights[self.offset - size:self.offset]

    def reset(self):
        self.offset = 4


def normalize(image):
    """utility/utils.py:150-153."""
    return image / 255.


def interval_overlap(interval_a, interval_b):
    x1, x2 = interval_a
    x3, x4 = interval_b
    if x3 < x1:
        return 0 if x4 < x1 else min(x2, x4) - x1
    return 0 if x2 < x3 else min(x2, x4) - x3


def bbox_iou(box1, box2):
    """utility/utils.py:155-173."""
    x1_min, x1_max = box1.x - box1.w / 2, box1.x + box1.w / 2
    x2_min, x2_max = box2.x - box2.w / 2, box2.x + box2.w / 2
    y1_min, y1_max = box1.y - box1.h / 2, box1.y + box1.h / 2
    y2_min, y2_max = box2.y - box2.h / 2, box2.y + box2.h / 2
    intersect = interval_overlap([x1_min, x1_max], [x2_min, x2_max]) * interval_overlap([y1_min, y1_max], [y2_min, y2_max])
    union = box1.w * box1.h + box2.w * box2.h - intersect
    return float(intersect) / union


def sigmoid(x):
    return 1. / (1. + np.exp(-x))


def softmax(x, axis=-1, t=-100.):
    x = x - np.max(x)
    if np.min(x) < t:
        x = x / np.min(x) * t
    e_x = np.exp(x)
    return e_x / e_x.sum(axis, keepdims=True)


_ENGINE = None


def _decode_engine():
    """A detector-less context is enough for the decode kernels (they take sizes per call)."""
    global _ENGINE
    if _ENGINE is None:
        from ..engine import DetectorEngine
        _ENGINE = DetectorEngine(n_class=2, max_batch=1)
    return _ENGINE


def boxes_from_rows(rows: np.ndarray, nb_class: int) -> List[BoundBox]:
    """(n,8) device rows [x,y,w,h,conf,score,label,anchor] -> BoundBox list (classes = post-NMS scores)."""
    out = []
    for r in rows:
        cls = np.zeros(nb_class, np.float32)
        cls[int(r[6])] = r[5]
        b = BoundBox(np.float32(r[0]), np.float32(r[1]), np.float32(r[2]), np.float32(r[3]), np.float32(r[4]), cls)
        out.append(b)
    return out


def decode_netout(netout, obj_threshold, nms_threshold, anchors, nb_class, engine=None) -> List[BoundBox]:
    """Reference signature (utility/utils.py:208).  ``netout``: (grid_h, grid_w, nb_box, 5+nb_class) numpy
    array or CUDA tensor of raw conv_23 outputs.  Unlike the reference the input is not mutated."""
    import torch
    eng = engine or _decode_engine()
    if isinstance(netout, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(netout, dtype=np.float32)).to(eng.device)
    else:
        t = netout.to(device=eng.device, dtype=torch.float32).contiguous()
    if t.shape[-1] != 5 + nb_class:
        raise ValueError(f"netout last dim {t.shape[-1]} != 5 + nb_class ({5 + nb_class})")
    boxes, counts = eng.decode(t[None], float(obj_threshold), float(nms_threshold), list(anchors))
    from ..engine import rows_to_host
    return boxes_from_rows(rows_to_host(boxes, counts)[0], nb_class)


def generate_heatmap_feat(det_x, det_y, det_w, det_h, hmap_size=32):
    """utility/utils.py:53-58 on the device; returns a (hmap_size**2,) numpy array like the reference."""
    import torch
    eng = _decode_engine()
    xywh = torch.tensor([[det_x, det_y, det_w, det_h]], dtype=torch.float32, device=eng.device)
    return eng.heatmap_from_box(xywh, hmap_size)[0].cpu().numpy().astype(np.float64)


def generate_rectangle_from_heatmap(heat_map, thresh=0.75, hmap_size=32):
    """utility/utils.py:61-79 on the device -> (x1, y1, x2, y2)."""
    import torch
    eng = _decode_engine()
    h = torch.as_tensor(np.asarray(heat_map, dtype=np.float32).reshape(1, -1), device=eng.device)
    r = eng.box_from_heatmap(h, hmap_size, thresh)[0].cpu().numpy()
    return int(r[0]), int(r[1]), int(r[2]), int(r[3])


def overlap_score(y_true, y_pred):
    """utility/utils.py:82-101: IoU of two corner boxes (x1, y1, x2, y2) as the reference's tracking metric computes
    it (|dx*dy| products, no clamping of an empty intersection).  Host arithmetic in Python floats, like the
    reference: it runs once per evaluated frame, after the accelerated path."""
    x1 = max(y_true[0], y_pred[0])
    y1 = max(y_true[1], y_pred[1])
    x2 = min(y_true[2], y_pred[2])
    y2 = min(y_true[3], y_pred[3])
    intersection = float(abs((x1 - x2) * (y1 - y2)))
    union = float(abs((y_true[0] - y_true[2]) * (y_true[1] - y_true[3]))) + \
        float(abs((y_pred[0] - y_pred[2]) * (y_pred[1] - y_pred[3]))) - intersection
    return intersection / union


def average_overlap_score(y_true, y_pred):
    """utility/utils.py:103-110 (mean of overlap_score over the paired samples)."""
    score, total = 0.0, 0
    for i, (t, p) in enumerate(zip(y_true, y_pred)):
        score += overlap_score(t, p)
        total = i
    return score / (total + 1)


def draw_boxes(image, boxes, labels):
    """utility/utils.py:190-206 (host, OpenCV): out of the accelerated path, kept for predict()."""
    import cv2
    for box in boxes:
        xmin = int((box.x - box.w / 2) * image.shape[1])
        xmax = int((box.x + box.w / 2) * image.shape[1])
        ymin = int((box.y - box.h / 2) * image.shape[0])
        ymax = int((box.y + box.h / 2) * image.shape[0])
        cv2.rectangle(image, (xmin, ymin), (xmax, ymax), (0, 255, 0), 3)
        cv2.putText(image, labels[box.get_label()] + ' ' + str(box.get_score()), (xmin, ymin - 13),
                    cv2.FONT_HERSHEY_SIMPLEX, 1e-3 * image.shape[0], (0, 255, 0), 2)
    return image
