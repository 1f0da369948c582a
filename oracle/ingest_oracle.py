"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- CPU restatement of the frame ingest step.

The reference resizes every frame on the host with OpenCV before the forward pass:
``cv2.resize(image, (IMAGE_W, IMAGE_H))`` at ``models_detection/KerasYOLO.py:526`` and
``models_tracking/MultiObjDetTracker.py:300`` (default interpolation INTER_LINEAR, uint8 BGR input).
OpenCV is a third-party dependency of the reference that is not vendored (README.md:12-18, version unpinned); its
published algorithm for 8-bit INTER_LINEAR (modules/imgproc/src/resize.cpp: the coefficient set-up of cv::resize,
HResizeLinear<uchar,int,short,2048>, VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>) is restated below.

Pinned: ``oracle/make_golden.py`` runs the installed cv2 (4.13.0 in this container) on seeded images and commits its
outputs in ``tests/golden/resize_cases.npz``; ``tests/test_oracle_cpu.py`` checks this restatement against them
bit for bit, and against cv2 itself when it is importable.
"""
from __future__ import annotations

import numpy as np


def _coef(x: np.ndarray) -> np.ndarray:
    """saturate_cast<short>(float): cvRound (round half to even), clamp to int16."""
    return np.clip(np.rint(x.astype(np.float64)), -32768, 32767).astype(np.int32)


def resize_tables(src_h: int, src_w: int, dst_h: int, dst_w: int):
    """Per destination column (sx, sx1, a0, a1) and row (sy0, sy1, b0, b1); 2048 = 1.0 (INTER_RESIZE_COEF_BITS 11)."""
    scale_x = 1.0 / (dst_w / src_w)
    scale_y = 1.0 / (dst_h / src_h)
    fx = ((np.arange(dst_w, dtype=np.float64) + 0.5) * scale_x - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int32)
    fx = (fx - sx.astype(np.float32)).astype(np.float32)
    lo = sx < 0                                   # resize.cpp: "if( sx < 0 ) fx = 0, sx = 0"
    fx[lo] = 0
    sx[lo] = 0
    hi = sx >= src_w - 1                          # "if( sx >= ssize.width-1 ) fx = 0, sx = ssize.width-1"
    fx[hi] = 0
    sx[hi] = src_w - 1
    a0 = _coef((np.float32(1.0) - fx) * np.float32(2048))
    a1 = _coef(fx * np.float32(2048))
    sx1 = np.minimum(sx + 1, src_w - 1)
    fy = ((np.arange(dst_h, dtype=np.float64) + 0.5) * scale_y - 0.5).astype(np.float32)
    sy = np.floor(fy).astype(np.int32)
    fy = (fy - sy.astype(np.float32)).astype(np.float32)
    b0 = _coef((np.float32(1.0) - fy) * np.float32(2048))
    b1 = _coef(fy * np.float32(2048))
    sy0 = np.clip(sy, 0, src_h - 1)               # rows are clipped when fetched, the weights are kept
    sy1 = np.clip(sy + 1, 0, src_h - 1)
    return (sx, sx1, a0, a1), (sy0, sy1, b0, b1)


def resize_linear_u8(src: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """cv2.resize(src, (dst_w, dst_h)) for an (H, W, C) uint8 image, bit for bit."""
    src = np.asarray(src)
    assert src.dtype == np.uint8 and src.ndim == 3
    (sx, sx1, a0, a1), (sy0, sy1, b0, b1) = resize_tables(src.shape[0], src.shape[1], dst_h, dst_w)
    s = src.astype(np.int32)
    h = s[:, sx, :] * a0[None, :, None] + s[:, sx1, :] * a1[None, :, None]          # horizontal pass, int
    r0, r1 = h[sy0], h[sy1]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


RESIZE_CASES = [  # (seed, src_h, src_w, dst) -- small enough to commit, cover up/down-scaling and odd sizes
    (1, 37, 53, 64), (2, 120, 160, 96), (3, 64, 64, 64), (4, 50, 200, 96), (5, 97, 31, 64), (6, 240, 320, 128),
]


def resize_case(seed: int, h: int, w: int) -> np.ndarray:
    return np.random.default_rng(1000 + seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
