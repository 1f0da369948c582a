#!/usr/bin/env python
"""CPU emulation of per-layer operand demotion: for each conv layer alone, replace the
3-term fp16-pair product  x*w ~ xh*wh + xh*wl + xl*wh  by a cheaper scheme and measure the max logit error
against the exact fp64 forward.  Operand rounding only (fp64 accumulation): the hardware's accumulator truncation
comes on top and is measured on the GPU (profiles/r2_precision_budget.md).

schemes:  xh    = drop the x_lo term (activations rounded to fp16)
          wh    = drop the w_lo term (weights rounded to fp16)
          hh    = one term (both rounded)
          f8    = corrections computed from e4m3-rounded factors: xh*wh + q8(xh)*q8(wl) + q8(xl)*q8(wh)

usage: python tests/tools/precision_budget_cpu.py [B] [n_class]
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from object_tracking_b200 import weights as W  # noqa: E402
from oracle import yolo_oracle as Y  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
C = int(sys.argv[2]) if len(sys.argv) > 2 else 80


def r16(t):
    return t.to(torch.float16).to(torch.float64)


def r8(t):
    return t.to(torch.float32).to(torch.float8_e4m3fn).to(torch.float32).to(torch.float64)


def conv_scheme(x, ker, k, scheme):
    if scheme == "exact":
        return F.conv2d(x, ker, padding=k // 2)
    # per-output-channel power-of-two scale like pack_conv (keeps fp16 away from subnormals)
    m = ker.abs().amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-30)
    up = torch.exp2(8 - torch.ceil(torch.log2(m)))
    ks = ker * up
    xh, wh = r16(x), r16(ks)
    xl, wl = r16(x - xh), r16(ks - wh)
    p = lambda a, b: F.conv2d(a, b, padding=k // 2)
    if scheme == "xh":
        y = p(xh, wh) + p(xh, wl)
    elif scheme == "wh":
        y = p(xh, wh) + p(xl, wh)
    elif scheme == "hh":
        y = p(xh, wh)
    elif scheme == "f8":
        # fp8 operands need their own scaling: x_lo ~ 2^-12 |x| -> scale by 2^11 (exact); w_lo likewise
        y = p(xh, wh) + p(r8(xh), r8(wl * 2048.0)) / 2048.0 + p(r8(xl * 2048.0), r8(wh / 64.0)) * 64.0 / 2048.0
    elif scheme == "3":
        y = p(xh, wh) + p(xh, wl) + p(xl, wh)
    else:
        raise ValueError(scheme)
    return y / up.view(1, -1, 1, 1)


def forward(frames, w, plan):
    dt = torch.float64
    x = torch.from_numpy(frames / 255.).to(dt).permute(0, 3, 1, 2).contiguous()
    skip = None
    ker = lambda i: torch.from_numpy(w[f"kernel_{i}"]).to(dt).permute(3, 2, 0, 1).contiguous()
    with torch.no_grad():
        for idx, (k, _ci, _co, pool) in enumerate(Y._TRUNK, start=1):
            x = conv_scheme(x, ker(idx), k, plan.get(idx, "exact") if idx > 1 else "exact")
            x = Y._bn_leaky(x, w, idx, "darknet", 1e-3, dt)
            if idx == 13:
                skip = x
            if pool:
                x = F.max_pool2d(x, 2, 2)
        s = conv_scheme(skip, ker(21), 1, plan.get(21, "exact"))
        s = Y._bn_leaky(s, w, 21, "darknet", 1e-3, dt)
        s = Y.darknet_reorg(s, 2)
        x = torch.cat([s, x], dim=1)
        x = conv_scheme(x, ker(22), 3, plan.get(22, "exact"))
        x = Y._bn_leaky(x, w, 22, "darknet", 1e-3, dt)
        x = conv_scheme(x, ker(23), 1, plan.get(23, "exact")) + torch.from_numpy(w["bias_23"]).to(dt).view(1, -1, 1, 1)
    return x.numpy()


def main():
    torch.set_num_threads(os.cpu_count())
    w = W.synthetic_yolo_weights(C, seed=0)
    frames = np.random.default_rng(99).integers(0, 256, (B, 416, 416, 3), dtype=np.uint8)
    ref = forward(frames, w, {})
    print(f"B={B} C={C} max|logit|={np.abs(ref).max():.2f}")
    all3 = forward(frames, w, {i: "3" for i in range(2, 24)})
    print(f"all layers 3-term (operand floor): {np.abs(all3 - ref).max():.2e}")
    for sch in ("xh", "wh", "hh", "f8"):
        e = forward(frames, w, {i: sch for i in range(2, 24)})
        print(f"all layers {sch}: {np.abs(e - ref).max():.2e}")
    print("layer | xh | wh | hh | f8   (that layer alone, others exact)")
    for i in range(2, 24):
        row = []
        for sch in ("xh", "wh", "hh", "f8"):
            e = forward(frames, w, {i: sch})
            row.append(np.abs(e - ref).max())
        print(f"conv_{i:<2d} | " + " | ".join(f"{v:.2e}" for v in row), flush=True)


if __name__ == "__main__":
    main()
