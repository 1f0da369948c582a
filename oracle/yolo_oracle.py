"""CPU oracle of the YOLOv2 forward pass.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates models_detection/KerasYOLO.py:277-405 (Keras graph) and, in ``mode="darknet"``, the
layer semantics of the darknet C forward (darknet/src/convolutional_layer.c:445-485,
batchnorm_layer.c:135-155 -> blas.c:147-158 ``normalize_cpu``, activations.h:38 leaky,
maxpool_layer.c:79-114, reorg_layer.c:91-110 -> blas.c:9-30 ``reorg_cpu``, route_layer.c:74-87).

torch-CPU ``conv2d`` is the only non-numpy arithmetic; float64 is the primary oracle, float32
quantifies the fp32 noise floor.  Its darknet mode is pinned against ``oracle/_ref/libdarknet.so`` (the reference's C
library): ``oracle/make_golden.py darknet`` asserts the agreement when it writes ``tests/golden/darknet_416.npz`` and
``tests/test_oracle_cpu.py`` re-checks the oracle against those committed library outputs.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# conv_k table: (index, ksize, cin, cout, bn, pool, src) -- restated from KerasYOLO.py:277-400
_TRUNK = [
    (3, 3, 32, True), (3, 32, 64, True), (3, 64, 128, False), (1, 128, 64, False),
    (3, 64, 128, True), (3, 128, 256, False), (1, 256, 128, False), (3, 128, 256, True),
    (3, 256, 512, False), (1, 512, 256, False), (3, 256, 512, False), (1, 512, 256, False),
    (3, 256, 512, True), (3, 512, 1024, False), (1, 1024, 512, False), (3, 512, 1024, False),
    (1, 1024, 512, False), (3, 512, 1024, False), (3, 1024, 1024, False), (3, 1024, 1024, False),
]


def normalize(image: np.ndarray) -> np.ndarray:
    """utility/utils.py:150-153: ``image / 255.`` (numpy true division -> float64)."""
    return image / 255.


def space_to_depth_x2(x_nhwc: torch.Tensor) -> torch.Tensor:
    """tf.space_to_depth(block_size=2), NHWC (KerasYOLO.py:241-242):
    out[b,i,j,(di*2+dj)*C + c] = in[b, 2i+di, 2j+dj, c]."""
    b, h, w, c = x_nhwc.shape
    x = x_nhwc.reshape(b, h // 2, 2, w // 2, 2, c)        # b i di j dj c
    x = x.permute(0, 1, 3, 2, 4, 5)                        # b i j di dj c
    return x.reshape(b, h // 2, w // 2, 4 * c)


def darknet_reorg(x_nchw: torch.Tensor, stride: int = 2) -> torch.Tensor:
    """darknet/src/blas.c:9-30 reorg_cpu(..., forward=0) as called by reorg_layer.c:108:
    out_flat[i + w*(j + h*k)] = in_flat[w2 + w*s*(h2 + h*s*c2)], a permutation of the flat CHW
    buffer that is then *viewed* as (c*s*s, h/s, w/s)."""
    b, c, h, w = x_nchw.shape
    out_c = c // (stride * stride)
    k = torch.arange(c).view(c, 1, 1)
    j = torch.arange(h).view(1, h, 1)
    i = torch.arange(w).view(1, 1, w)
    c2 = k % out_c
    offset = k // out_c
    w2 = i * stride + offset % stride
    h2 = j * stride + offset // stride
    src = (w2 + w * stride * (h2 + h * stride * c2)).reshape(-1)       # index into flat input
    flat = x_nchw.reshape(b, -1)
    out = flat[:, src]
    return out.reshape(b, c * stride * stride, h // stride, w // stride)


def _bn_leaky(x: torch.Tensor, w: Dict[str, np.ndarray], i: int, mode: str, eps: float, dt) -> torch.Tensor:
    g = torch.from_numpy(np.asarray(w[f"gamma_{i}"])).to(dt).view(1, -1, 1, 1)
    b = torch.from_numpy(np.asarray(w[f"beta_{i}"])).to(dt).view(1, -1, 1, 1)
    m = torch.from_numpy(np.asarray(w[f"mean_{i}"])).to(dt).view(1, -1, 1, 1)
    v = torch.from_numpy(np.asarray(w[f"var_{i}"])).to(dt).view(1, -1, 1, 1)
    if mode == "darknet":
        x = (x - m) / (torch.sqrt(v) + 1e-6) * g + b            # blas.c:156 then scale_bias, add_bias
    else:
        x = g * (x - m) / torch.sqrt(v + eps) + b               # keras BN inference, eps=1e-3
    return torch.where(x > 0, x, 0.1 * x)                        # LeakyReLU(0.1) / activations.h:38


def _conv(x: torch.Tensor, w: Dict[str, np.ndarray], i: int, k: int, dt) -> torch.Tensor:
    ker = torch.from_numpy(np.asarray(w[f"kernel_{i}"])).to(dt).permute(3, 2, 0, 1).contiguous()
    return F.conv2d(x, ker, padding=k // 2)                     # 'same', stride 1, cross-correlation


def yolo_forward(frames_nhwc: np.ndarray, w: Dict[str, np.ndarray], n_class: int,
                 dtype=np.float64, mode: str = "keras", bn_eps: float = 1e-3,
                 want: Optional[List[str]] = None) -> Dict[str, np.ndarray]:
    """Run conv_1..conv_23 on normalised frames (B,H,W,3) float.

    Returns NHWC arrays: ``logits`` (B,G,G,5,5+C) [conv_23 + Reshape, KerasYOLO.py:399-400],
    ``feat`` (B,G,G,1024) [layer 'conv_feat', :396], plus any ``conv_k`` / ``norm_k`` (post
    LeakyReLU, pre-pool) named in ``want``.
    """
    dt = torch.float64 if dtype == np.float64 else torch.float32
    want = set(want or [])
    out: Dict[str, np.ndarray] = {}
    x = torch.from_numpy(np.ascontiguousarray(frames_nhwc)).to(dt).permute(0, 3, 1, 2).contiguous()
    skip = None

    def keep(name: str, t: torch.Tensor) -> None:
        if name in want:
            out[name] = t.permute(0, 2, 3, 1).contiguous().numpy()

    with torch.no_grad():
        for idx, (k, _ci, _co, pool) in enumerate(_TRUNK, start=1):
            x = _conv(x, w, idx, k, dt)
            keep(f"conv_{idx}", x)
            x = _bn_leaky(x, w, idx, mode, bn_eps, dt)
            keep(f"norm_{idx}", x)
            if idx == 13:
                skip = x
            if pool:
                x = F.max_pool2d(x, 2, 2)
        s = _conv(skip, w, 21, 1, dt)
        keep("conv_21", s)
        s = _bn_leaky(s, w, 21, mode, bn_eps, dt)
        keep("norm_21", s)
        if mode == "darknet":
            s = darknet_reorg(s, 2)
        else:
            s = space_to_depth_x2(s.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
        x = torch.cat([s, x], dim=1)                             # concatenate([skip, x]) :391
        keep("concat", x)
        x = _conv(x, w, 22, 3, dt)
        keep("conv_22", x)
        x = _bn_leaky(x, w, 22, mode, bn_eps, dt)
        feat = x
        x = _conv(x, w, 23, 1, dt) + torch.from_numpy(np.asarray(w["bias_23"])).to(dt).view(1, -1, 1, 1)
    b, _, g_h, g_w = x.shape
    out["feat"] = feat.permute(0, 2, 3, 1).contiguous().numpy()
    out["logits"] = x.permute(0, 2, 3, 1).contiguous().numpy().reshape(b, g_h, g_w, 5, 5 + n_class)
    return out
