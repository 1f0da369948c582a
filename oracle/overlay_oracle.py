"""CPU oracle of the rectangle overlay (utility/utils.py:190-206 draw_boxes -> cv2.rectangle(..., thickness 3)).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

OpenCV is an un-vendored dependency of the reference; its thickness-3 rectangle is restated as a pixel-set rule and
pinned against the installed cv2 itself by tests/test_oracle_cpu.py: a thick segment covers
{ max(0, a - u, u - b) + |v - c| <= 2 } (u along the segment [a, b], v across it at c), a rectangle is the union of
its four edges.  Box corners: int((x -/+ w/2) * W), int((y -/+ h/2) * H) in float32 (numpy >= 2 arithmetic of the
reference on float32 boxes)."""
import numpy as np


def rectangle3(img: np.ndarray, xa: int, ya: int, xb: int, yb: int, color) -> None:
    H, W = img.shape[:2]
    if xa > xb:
        xa, xb = xb, xa
    if ya > yb:
        ya, yb = yb, ya
    yy, xx = np.mgrid[0:H, 0:W]
    ox = np.maximum(0, np.maximum(xa - xx, xx - xb))
    oy = np.maximum(0, np.maximum(ya - yy, yy - yb))
    m = np.zeros((H, W), bool)
    for c in (ya, yb):
        m |= ox + np.abs(yy - c) <= 2
    for c in (xa, xb):
        m |= oy + np.abs(xx - c) <= 2
    img[m] = color


def box_corners(row, W: int, H: int):
    x, y, w, h = (np.float32(v) for v in row[:4])
    two = np.float32(2)
    return (int((x - w / two) * np.float32(W)), int((y - h / two) * np.float32(H)),
            int((x + w / two) * np.float32(W)), int((y + h / two) * np.float32(H)))


def draw_boxes(img: np.ndarray, rows: np.ndarray, color=(0, 255, 0)) -> np.ndarray:
    for r in rows:
        xa, ya, xb, yb = box_corners(r, img.shape[1], img.shape[0])
        rectangle3(img, xa, ya, xb, yb, color)
    return img
