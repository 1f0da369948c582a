"""Margin-aware comparison of decoded boxes (test helper).

The CUDA logits differ from the fp64 oracle's by a few 1e-4, so a decision that sits on a threshold in the oracle
(class score == obj_threshold, IoU == nms_threshold) may legitimately fall on the other side on the device.  A blind
"skip the frame when the counts differ" hides real bugs; this helper instead matches boxes by identity
(anchor id + label, or class + position), compares every matched pair, and demands that every unmatched box is
EXPLAINED by a near-threshold decision visible in the data: its score is within `score_margin` of the threshold, or it
overlaps a same-class box with an IoU within `iou_margin` of the NMS threshold (suppressed on one side only), or the
box that suppresses it on the other side is itself unmatched (a flip that cascades).
"""
import numpy as np


def _iou_centre(a, b):
    def ov(x1, w1, x2, w2):
        return min(x1 + w1 / 2, x2 + w2 / 2) - max(x1 - w1 / 2, x2 - w2 / 2)
    w, h = ov(a[0], a[2], b[0], b[2]), ov(a[1], a[3], b[1], b[3])
    inter = 0.0 if (w <= 0 or h <= 0) else w * h
    return inter / (a[2] * a[3] + b[2] * b[3] - inter)


def compare_rows(got, ref, score_col=5, label_col=6, id_col=7, coord_tol=1e-3, score_thr=0.5, nms_thr=0.45,
                 score_margin=5e-3, iou_margin=2e-2, by_position=False, coord_scale=1.0):
    """got / ref: (n,8) rows [x,y,w,h, conf|objectness, score, label, anchor id].  Returns a dict with the worst
    coordinate error over matched rows, the matched / unmatched counts and the list of UNEXPLAINED unmatched rows
    (must be empty).  by_position: match by (label, nearest centre) instead of (anchor id, label)."""
    got, ref = np.asarray(got, np.float64).reshape(-1, 8), np.asarray(ref, np.float64).reshape(-1, 8)
    pairs, used = [], set()
    for i, g in enumerate(got):
        best = None
        for j, r in enumerate(ref):
            if j in used or int(r[label_col]) != int(g[label_col]):
                continue
            if by_position:
                d = max(abs(g[0] - r[0]), abs(g[1] - r[1]))
                if d < 0.02 * coord_scale and (best is None or d < best[1]):
                    best = (j, d)
            elif int(r[id_col]) == int(g[id_col]):
                best = (j, 0.0)
                break
        if best is not None:
            used.add(best[0])
            pairs.append((i, best[0]))
    worst = max([np.abs(got[i, :4] - ref[j, :4]).max() / coord_scale for i, j in pairs], default=0.0)
    worst_score = max([abs(got[i, score_col] - ref[j, score_col]) for i, j in pairs], default=0.0)
    un_g = [i for i in range(len(got)) if i not in {p[0] for p in pairs}]
    un_r = [j for j in range(len(ref)) if j not in used]
    every = [got[i] for i in range(len(got))] + [ref[j] for j in un_r]
    unexplained = []
    for row in [got[i] for i in un_g] + [ref[j] for j in un_r]:
        if row[score_col] <= score_thr + score_margin:
            continue                                            # score on the threshold
        near = False
        for other in every:
            if other is row or int(other[label_col]) != int(row[label_col]):
                continue
            if abs(_iou_centre(row, other) - nms_thr) <= iou_margin:
                near = True                                     # suppression decision on the NMS threshold
                break
        if not near:
            unexplained.append(row.tolist())
    return {"worst": worst, "worst_score": worst_score, "matched": len(pairs), "unmatched": len(un_g) + len(un_r),
            "unexplained": unexplained}
