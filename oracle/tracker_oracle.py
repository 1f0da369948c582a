"""CPU oracle of the recurrent tracker steps.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates
  models_tracking/TinyTracker.py:25-41         pool -> concat[fv, bbox4] -> LSTM(512, impl=2) -> Dense(4, sigmoid)
  models_tracking/TinyHeatmapTracker.py:26-48  same with a 32x32 heat-map in and Dense(1024, sigmoid) out
  utility/utils.py:53-58 / :61-79              generate_heatmap_feat / generate_rectangle_from_heatmap
  models_tracking/MultiObjDetTracker.py:160-189 concat[conv_23 logits, conv_feat] -> ConvLSTM2D(512,3x3,same)
                                                -> Conv2D(A*(5+C),1x1,+bias) -> Reshape
  utility/preprocessing.py:418-456             how the per-frame tracker inputs are formed

Keras itself is not vendored in the reference and not installed here; the gate equations are the
published Keras 2.0-2.2 ones (recurrent_activation = hard_sigmoid, gate order i,f,c,o, zero
initial state) -- parity for these rows is "unpinned" (DESIGN.md).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np


def hard_sigmoid(x):
    return np.clip(0.2 * x + 0.5, 0.0, 1.0)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def pool_features(feat_hwc: np.ndarray, pool: str, ref_layout_bug: bool = False) -> np.ndarray:
    """TinyTracker.py:29-33.  feat (H,W,C).  'Global' -> (C,) per-channel max;
    'Max' -> MaxPooling2D(4,4)/4 valid + Flatten -> ((H//4)*(W//4)*C,) in (h,w,c) order.
    ref_layout_bug: reproduce preprocessing.py:419 -- the detector's CHW buffer is *reshaped* as
    (H,W,C) without a transpose."""
    h, w, c = feat_hwc.shape
    if ref_layout_bug:
        feat_hwc = np.ascontiguousarray(np.transpose(feat_hwc, (2, 0, 1))).reshape(h, w, c)
    if pool == "Global":
        return feat_hwc.max(axis=(0, 1))
    if pool == "Max":
        ph, pw = h // 4, w // 4
        x = feat_hwc[:ph * 4, :pw * 4].reshape(ph, 4, pw, 4, c).max(axis=(1, 3))
        return x.reshape(-1)
    raise ValueError(pool)


def lstm_step(x, h, c, w: Dict[str, np.ndarray], recurrent_activation=hard_sigmoid):
    """One Keras-2 LSTM step (implementation=2): z = xW + hU + b, gates i,f,c,o."""
    u = h.shape[-1]
    z = x @ w["kernel"].astype(x.dtype) + h @ w["recurrent_kernel"].astype(x.dtype) + w["bias"].astype(x.dtype)
    i = recurrent_activation(z[..., :u])
    f = recurrent_activation(z[..., u:2 * u])
    c_new = f * c + i * np.tanh(z[..., 2 * u:3 * u])
    o = recurrent_activation(z[..., 3 * u:])
    h_new = o * np.tanh(c_new)
    return h_new, c_new


def tracker_step(fv, det, h, c, w, recurrent_activation=hard_sigmoid):
    """fv (S,F) pooled feature, det (S,4 | S,1024) -> y = sigmoid(Dense(h')), new (h,c)."""
    x = np.concatenate([fv, det], axis=-1)
    h, c = lstm_step(x, h, c, w, recurrent_activation)
    y = sigmoid(h @ w["dense_kernel"].astype(x.dtype) + w["dense_bias"].astype(x.dtype))
    return y, h, c


def generate_heatmap_feat(det_x, det_y, det_w, det_h, hmap_size=32):
    """utils.py:53-58 (note: x,y are the box's top-left, python int() truncation, +1 inclusive)."""
    heat = np.zeros((hmap_size, hmap_size))
    sx, sy, sh, sw = int(det_x * hmap_size), int(det_y * hmap_size), int(det_h * hmap_size), int(det_w * hmap_size)
    heat[sy:(sy + sh + 1), sx:(sx + sw + 1)] = 1.0
    return heat.reshape(-1)


def generate_rectangle_from_heatmap(heat_map, thresh=0.75, hmap_size=32):
    """utils.py:61-79 -> (x1, y1, x2, y2) cell indices; (size,size,-1,-1) if nothing >= thresh."""
    ys, xs = np.nonzero(np.asarray(heat_map).reshape(hmap_size, hmap_size) >= thresh)
    if ys.size == 0:
        return hmap_size, hmap_size, -1, -1
    return int(xs.min()), int(ys.min()), int(xs.max()), int(ys.max())


def detection_to_tracker_input(obj_det, frame_w, frame_h, heatmap_size: Optional[int] = None):
    """preprocessing.py:434-456: first (highest-prob) detection -> [cx/w, cy/h, bw/w, bh/h]
    (zeros if none), or its heat-map."""
    if len(obj_det) != 0:
        cx, cy, bw, bh = obj_det[0][2]
        v = (cx / frame_w, cy / frame_h, bw / frame_w, bh / frame_h)
    else:
        v = (0, 0, 0, 0)
    if heatmap_size is None:
        return np.array(v, dtype="float32")
    return generate_heatmap_feat(v[0] - v[2] / 2.0, v[1] - v[3] / 2.0, v[2], v[3], hmap_size=heatmap_size)


def _conv_same(x_hwc: np.ndarray, k_hwio: np.ndarray) -> np.ndarray:
    import torch
    import torch.nn.functional as F
    x = torch.from_numpy(np.ascontiguousarray(x_hwc)).permute(2, 0, 1)[None]
    k = torch.from_numpy(np.ascontiguousarray(k_hwio)).to(x.dtype).permute(3, 2, 0, 1).contiguous()
    return F.conv2d(x, k, padding=k_hwio.shape[0] // 2)[0].permute(1, 2, 0).numpy()


def convlstm_step(z_hwc, h, c, w: Dict[str, np.ndarray], recurrent_activation=hard_sigmoid):
    """Keras-2 ConvLSTM2D(units,3x3,'same') step: gates = conv(z,W)+b + conv(h,U); i,f,c,o."""
    u = h.shape[-1]
    g = _conv_same(z_hwc, w["kernel"]) + w["bias"].astype(z_hwc.dtype) + _conv_same(h, w["recurrent_kernel"])
    i = recurrent_activation(g[..., :u])
    f = recurrent_activation(g[..., u:2 * u])
    c_new = f * c + i * np.tanh(g[..., 2 * u:3 * u])
    o = recurrent_activation(g[..., 3 * u:])
    h_new = o * np.tanh(c_new)
    return h_new, c_new


def multiobj_step(logits_hwc, feat_hwc, h, c, w, recurrent_activation=hard_sigmoid):
    """MultiObjDetTracker.py:175-183 for one frame: z = concat[x_bbox, x_vis] -> ConvLSTM -> 1x1 head.
    Returns (tracker logits (G,G,A*(5+C)), h, c)."""
    z = np.concatenate([logits_hwc, feat_hwc], axis=-1)
    h, c = convlstm_step(z, h, c, w, recurrent_activation)
    out = _conv_same(h, w["head_kernel"]) + w["head_bias"].astype(h.dtype)
    return out, h, c


def make_cases() -> Dict[str, np.ndarray]:
    """Seeded regression vectors for the tracker kernels (fp64 oracle outputs)."""
    import importlib
    W = importlib.import_module("object-tracking_b200.weights")
    out: Dict[str, np.ndarray] = {}
    rng = np.random.default_rng(77)
    # TinyTracker Global: 1024 + 4 -> 512 -> 4 ; 6 steps, 3 streams, reset every 4
    for tag, n_det, n_out in (("tiny", 4, 4), ("heat", 1024, 1024)):
        w = W.synthetic_lstm_weights(1024 + n_det, 512, n_out, seed=11)
        S, T = 3, 6
        fv = np.abs(rng.standard_normal((T, S, 1024)))
        det = rng.uniform(0, 1, (T, S, n_det)) if n_det == 4 else (rng.uniform(0, 1, (T, S, n_det)) > 0.8) * 1.0
        h = np.zeros((S, 512)); c = np.zeros((S, 512))
        ys = []
        for t in range(T):
            if t % 4 == 0:
                h[:] = 0; c[:] = 0
            y, h, c = tracker_step(fv[t], det[t], h, c, {k: v.astype(np.float64) for k, v in w.items()})
            ys.append(y)
        out[f"{tag}_fv"] = fv.astype(np.float32)
        out[f"{tag}_det"] = det.astype(np.float32)
        out[f"{tag}_y"] = np.stack(ys)
        out[f"{tag}_h"] = h
        out[f"{tag}_c"] = c
    return out
