"""GPU parity tests added in round 2 (run with -m gpu on a B200): the holes VERDICT r1 listed.

 * the detection choice of utility/preprocessing.py:434-456 (b2t_select_detection) on both branches, directly and
   through TinyTracker with the planted synthetic detector;
 * the benched configuration itself: 36 frames, C=80, darknet semantics, against darknet_oracle;
 * the heat-map kernels against outputs of the REFERENCE functions (tests/golden/heatmap_cases.npz);
 * MultiObjDetTracker at its real size (ConvLSTM2D(512), C=20, T=4), multi-stream windows, step();
 * config 5: 608x608, batch 8, TinyHeatmapTracker with pool="Max".
Tolerances: logits vs the fp64 oracle < 1e-3 absolute (north_star's bbox bar: dw = w * dlogit); discrete outputs
bit-exact on the same logits; boxes across the whole chain compared with tests/parity_util.compare_rows."""
import os

import numpy as np
import pytest
import torch

from oracle import darknet_oracle, decode_oracle, tracker_oracle, yolo_oracle
from oracle.cases import heatmap_case_inputs
from object_tracking_b200 import weights as W
from parity_util import compare_rows

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

TRACKER_CFG = {"model_detector": {"name": "YOLO", "config_file": "cfg/yolov2.cfg", "meta_file": "cfg/coco.data",
                                  "weights_file": "none.weights", "fv_layer": 25, "nms": 0.45, "thresh": 0.5, "hier_thresh": 0.5},
               "model_tracker": {"name": "TinyTracker", "lstm_units": 512, "sequence_length": 4, "heatmap_size": 32},
               "train": {"cpu_only": 0, "dgpu_id": 0, "tgpu_id": 0, "pool": "Global", "batch_size": 4, "max_epochs": 0,
                         "tensorboard_dir": "logs/", "saved_model_dir": "models/", "classes": ["Person", "Car"]}}


def _engine(**kw):
    from object_tracking_b200.engine import DetectorEngine
    return DetectorEngine(**kw)


# ------------------------------------------------------------------------------------------------ a15
def _select_case(rng, B, max_dets, n_class):
    """Detection rows as b2t_region_detect emits them (sorted by -prob) with crafted class patterns."""
    dets = np.zeros((B, max_dets, 8), np.float32)
    counts = rng.integers(0, max_dets + 1, B).astype(np.int32)
    counts[0] = 0                                              # empty frame
    for b in range(B):
        n = int(counts[b])
        prob = np.sort(rng.uniform(0.5, 1.0, n))[::-1]
        if n >= 3 and b % 3 == 0:
            prob[:3] = prob[0]                                 # ties: the first row in the given order wins
        dets[b, :n, 0] = rng.uniform(0, 640, n); dets[b, :n, 1] = rng.uniform(0, 480, n)
        dets[b, :n, 2] = rng.uniform(1, 300, n); dets[b, :n, 3] = rng.uniform(1, 300, n)
        dets[b, :n, 4] = rng.uniform(0.5, 1, n); dets[b, :n, 5] = prob
        dets[b, :n, 6] = rng.integers(0, n_class, n)
        dets[b, :n, 7] = rng.integers(0, 845, n)
        dets[b, n:] = rng.uniform(-5, 5, (max_dets - n, 8))    # stale rows past the count must be ignored
    if B > 1 and counts[1] > 0:
        dets[1, :counts[1], 6] = 5                             # frame 1: no allowed class at all
    return dets, counts


@pytest.mark.parametrize("masked", [True, False])
@pytest.mark.parametrize("heat", [0, 32])
def test_select_detection_against_oracle(masked, heat):
    """b2t_select_detection == tracker_oracle.detection_to_tracker_input on the class-filtered list
    (preprocessing.py:434-456, YOLO.py:177-180): masked and unmasked, empty, no allowed class, ties, non-square
    frame, bbox and heat-map modes."""
    eng = _engine(n_class=2, max_batch=1)
    rng = np.random.default_rng(17 + heat)
    B, md, nc, fw, fh = 24, 12, 8, 640, 480
    dets, counts = _select_case(rng, B, md, nc)
    allowed = [0, 2]
    mask = torch.zeros(nc, dtype=torch.uint8)
    mask[allowed] = 1
    det_in, hm, chosen = eng.select_detection(torch.from_numpy(dets).cuda(), torch.from_numpy(counts).cuda(), fw, fh,
                                              mask.cuda() if masked else None, heat)
    det_in, chosen = det_in.cpu().numpy(), chosen.cpu().numpy()
    n_pos = 0
    for b in range(B):
        rows = dets[b, :counts[b]]
        lst = [("c%d" % int(r[6]), float(r[5]), tuple(float(v) for v in r[:4])) for r in rows
               if (not masked) or int(r[6]) in allowed]
        first = next((i for i, r in enumerate(rows) if (not masked) or int(r[6]) in allowed), -1)
        assert int(chosen[b]) == first, b
        ref = tracker_oracle.detection_to_tracker_input(lst, fw, fh)
        assert np.array_equal(det_in[b], ref), (b, det_in[b], ref)
        n_pos += first >= 0
        if heat:
            ref_h = tracker_oracle.detection_to_tracker_input(lst, fw, fh, heatmap_size=heat)
            assert np.array_equal(hm.cpu().numpy()[b], ref_h.astype(np.float32)), b
    assert 3 <= n_pos < B                                       # both branches exercised
    if masked:
        assert chosen[1] == -1 and not det_in[1].any()


def test_tiny_tracker_positive_detection_branch():
    """TinyTracker.track_windows against the full oracle chain with the planted detector: most frames carry a
    person / car detection, so the class filter, the top-probability pick and the /frame_w,h normalisation feed
    NON-ZERO rows to the LSTM (they never had)."""
    from object_tracking_b200.models_tracking.TinyTracker import TinyTracker
    trk = TinyTracker(TRACKER_CFG, max_streams=2)
    assert trk.model_detector.synthetic_weights
    frames = np.random.default_rng(11).integers(0, 256, (2, 4, 416, 416, 3), dtype=np.uint8)
    y = trk.track_windows(torch.from_numpy(frames).cuda(), graph=False).cpu().numpy()
    eng = trk.model_detector.engine
    _, det_in, _, chosen = trk._decode_and_pool(8, 416, 416)
    det_in, chosen = det_in.cpu().numpy(), chosen.cpu().numpy()
    w = W.synthetic_detector_weights(80, seed=0)
    wl = {k: v.astype(np.float64) for k, v in W.synthetic_lstm_weights(1028, 512, 4, seed=1).items()}
    o = yolo_oracle.yolo_forward(yolo_oracle.normalize(frames.reshape(8, 416, 416, 3)), w, 80, dtype=np.float64,
                                 mode="darknet", want=["norm_20"])
    names = trk.model_detector.names
    n_with_det = 0
    for s in range(2):
        h = np.zeros((1, 512)); c = np.zeros((1, 512))
        for t in range(4):
            i = s * 4 + t
            logits = np.transpose(o["logits"][i].reshape(13, 13, -1), (2, 0, 1)).astype(np.float32)
            region = darknet_oracle.region_forward(logits, 80)
            boxes, obj, prob = darknet_oracle.detect(region, 416, 416, 416, 416, 0.5, 0.45, 80)
            lst = [d for d in darknet_oracle.yolo_detect_list(boxes, obj, prob, names) if d[0] in ("person", "car")]
            ref_in = tracker_oracle.detection_to_tracker_input(lst, 416, 416)
            # the choice is only comparable when the oracle's top allowed detection is not on the 0.5 threshold and
            # is not contested by a runner-up within the logit noise
            clear = (not lst and int(chosen[i]) < 0) or \
                    (lst and lst[0][1] > 0.505 and (len(lst) < 2 or lst[0][1] - lst[1][1] > 2e-3))
            if clear:
                assert (int(chosen[i]) >= 0) == bool(lst), i
                assert np.abs(det_in[i] - ref_in).max() < 1e-3, (i, det_in[i], ref_in)
                n_with_det += bool(lst)
            fv = o["norm_20"][i].max(axis=(0, 1))[None]
            # feed the oracle LSTM what the device chose, so a legitimate threshold flip does not derail the rest
            yy, h, c = tracker_oracle.tracker_step(fv, det_in[i][None].astype(np.float64), h, c, wl)
            assert np.abs(y[s, t] - yy[0]).max() < 1e-3, (s, t)
    assert n_with_det >= 3, n_with_det
    assert (chosen >= 0).sum() >= 3
    # online stepping of two interleaved streams (state slot per stream) reproduces the windows
    trk.reset()
    for t in range(4):
        for s in (1, 0):
            ys = trk.step(frames[s, t], stream=s)
            assert np.abs(ys - y[s, t]).max() < 1e-5, (s, t)
    # det_bbox replaces the detector's choice (fifth step of stream 0 = first step of a new window)
    y5 = trk.step(frames[0, 0], det_bbox=(208.0, 104.0, 100.0, 50.0), stream=0)
    ref_in = np.array([[208 / 416, 104 / 416, 100 / 416, 50 / 416]])
    yy, _, _ = tracker_oracle.tracker_step(o["norm_20"][0].max(axis=(0, 1))[None], ref_in, np.zeros((1, 512)),
                                           np.zeros((1, 512)), wl)
    assert np.abs(y5 - yy[0]).max() < 1e-3
    with pytest.raises(ValueError):
        trk.step(frames[0, 0], stream=2)


# ------------------------------------------------------------------------------------------------ benched config
def test_benchmark_config_c80_darknet_batch36():
    """bench.py's own configuration -- 36 frames per step, C=80, darknet semantics, planted detector -- against the
    fp64 oracle forward (darknet mode) and darknet_oracle's region layer + do_nms_obj."""
    C, B = 80, 36
    w = W.synthetic_detector_weights(C, seed=0)
    frames = np.random.default_rng(1234).integers(0, 256, (B, 416, 416, 3), dtype=np.uint8)
    eng = _engine(n_class=C, max_batch=B, semantics="darknet")
    eng.set_weights(w)
    eng.finalize()
    logits = eng.forward(torch.from_numpy(frames).cuda())
    dets, counts = eng.region_detect(logits, 0.5, 0.45, 416, 416)
    got = logits.cpu().numpy()
    o = yolo_oracle.yolo_forward(yolo_oracle.normalize(frames), w, C, dtype=np.float64, mode="darknet", want=["norm_20"])
    # Error bars per entry type.  The bbox bar of north_star (coordinates within 1e-3) is on t_x, t_y, t_w, t_h:
    # d(x, y) <= |dt| / (4 * 13), d(w, h) = (w, h) * |dt|.  Objectness and class logits only enter through sigmoid /
    # softmax (slope <= 1/4), their bar is on the resulting scores (checked below: worst_score < 1e-3); with the planted
    # class bias they reach |logit| = 25, and the error of a logit is relative to the size of its sum (measured: 1.3e-3
    # absolute = 5e-5 relative at |25|).
    d = np.abs(got - o["logits"])
    assert d[..., :4].max() < 5e-4, d[..., :4].max()
    assert d.max() < 1e-4 * np.abs(o["logits"]).max(), (d.max(), np.abs(o["logits"]).max())
    fv = eng.pool_features("norm_20", B, "Global").cpu().numpy()
    assert np.abs(fv - o["norm_20"].max(axis=(1, 2))).max() < 3e-3
    dets, counts = dets.cpu().numpy(), counts.cpu().numpy()
    worst, matched, unmatched, total = 0.0, 0, 0, 0
    for b in range(B):
        lg = np.transpose(o["logits"][b].reshape(13, 13, -1), (2, 0, 1)).astype(np.float32)
        region = darknet_oracle.region_forward(lg, C)
        boxes, obj, prob = darknet_oracle.detect(region, 416, 416, 416, 416, 0.5, 0.45, C)
        ref = [(bx, ob, pr.max(), pr.argmax()) for bx, ob, pr in zip(boxes, obj, prob) if pr.max() > 0]
        ref = np.array([[*bx, ob, p, c, -1] for bx, ob, p, c in ref], np.float64).reshape(-1, 8)
        r = compare_rows(dets[b, :counts[b]], ref, by_position=True, coord_scale=416.0, score_margin=5e-3)
        assert not r["unexplained"], (b, r["unexplained"])
        assert r["worst_score"] < 1e-3
        worst = max(worst, r["worst"])
        matched += r["matched"]; unmatched += r["unmatched"]; total += len(ref)
    assert worst < 1e-3, worst                       # box coordinates relative to the frame (north_star bar)
    assert matched >= 60 and unmatched <= max(2, total // 10), (matched, unmatched, total)


# ------------------------------------------------------------------------------------------------ a12
def test_heatmap_kernels_match_reference_outputs():
    """Outputs of the REFERENCE's generate_heatmap_feat / generate_rectangle_from_heatmap (utils.py:53-79), committed
    by oracle/make_golden.py -- not of our restatement."""
    eng = _engine(n_class=2, max_batch=1)
    z = np.load(os.path.join(GOLD, "heatmap_cases.npz"))
    n = int(z["n"])
    xywh, heat = heatmap_case_inputs(n, int(z["seed"]))
    feats = np.unpackbits(z["feat_bits"], axis=1)[:, :1024].astype(np.float32)
    got = eng.heatmap_from_box(torch.from_numpy(xywh.astype(np.float32)).cuda(), 32).cpu().numpy()
    assert np.array_equal(got, feats)
    rect = eng.box_from_heatmap(torch.from_numpy(heat.astype(np.float32)).cuda(), 32, 0.75).cpu().numpy()
    assert np.array_equal(rect, z["rect"])
    rect2 = eng.box_from_heatmap(torch.from_numpy(feats).cuda(), 32, 0.75).cpu().numpy()
    assert np.array_equal(rect2, z["rect_of_feat"])


# ------------------------------------------------------------------------------------------------ a13 / a14 (C3)
@pytest.fixture(scope="module")
def c3():
    """MultiObjDetTracker at its real size: C=20, ConvLSTM2D(512), 2 streams x 4 frames + the fp64 oracle chain."""
    from object_tracking_b200.models_tracking.MultiObjDetTracker import MultiObjDetTracker
    C, U, S, T = 20, 512, 2, 4
    labels = [str(i) for i in range(C)]
    wd = W.synthetic_yolo_weights(C, seed=0)
    wl = W.synthetic_multiobj_weights(C, U, seed=2)
    mt = MultiObjDetTracker({"LABELS": labels}, detector_weights=wd, tracker_weights=wl, max_streams=S)
    frames = np.random.default_rng(99).integers(0, 256, (S, T, 416, 416, 3), dtype=np.uint8)
    o = yolo_oracle.yolo_forward(yolo_oracle.normalize(frames.reshape(S * T, 416, 416, 3)), wd, C, dtype=np.float64)
    wl64 = {k: v.astype(np.float64) for k, v in wl.items()}
    refs = np.zeros((S, T, 13, 13, 5 * (5 + C)))
    for s in range(S):
        h = np.zeros((13, 13, U)); c = np.zeros((13, 13, U))
        for t in range(T):
            i = s * T + t
            refs[s, t], h, c = tracker_oracle.multiobj_step(o["logits"][i].reshape(13, 13, -1), o["feat"][i], h, c, wl64)
    return mt, frames, o, refs, (C, U, S, T)


def test_multiobj_full_size_logits_and_decode(c3):
    mt, frames, o, refs, (C, U, S, T) = c3
    fr = torch.from_numpy(frames).cuda()
    trk_logits, boxes, counts = mt.track_windows(fr, reset=True, graph=False)
    lg = trk_logits.cpu().numpy().reshape(S, T, 13, 13, -1)
    err = np.abs(lg - refs).max()
    assert err < 1e-3, err                                      # |logit| up to ~12
    # decode of the device's own tracker logits: discrete outputs bit-exact, floats <= 2e-6
    rows, n = boxes.cpu().numpy(), counts.cpu().numpy()
    total = 0
    for i in range(S * T):
        ref = decode_oracle.boxes_to_array(decode_oracle.decode_netout(
            trk_logits[i].cpu().numpy(), 0.5, 0.45, W.ANCHORS, C))
        assert n[i] == len(ref), (i, n[i], len(ref))
        got = rows[i, :n[i]].astype(np.float64)
        assert np.array_equal(got[:, 6:8], ref[:, 6:8])
        assert np.abs(got[:, :6] - ref[:, :6]).max(initial=0) < 2e-6
        total += n[i]
    assert total > 100                                          # MOT17-like box counts, not an empty decode
    # the whole chain against the oracle chain (fp64 forward -> ConvLSTM -> head -> decode), margin-aware
    worst = 0.0
    for i in range(S * T):
        ref = decode_oracle.boxes_to_array(decode_oracle.decode_netout(
            refs.reshape(S * T, 13, 13, 5, 5 + C)[i].astype(np.float32), 0.5, 0.45, W.ANCHORS, C))
        r = compare_rows(rows[i, :n[i]], ref)
        assert not r["unexplained"], (i, r["unexplained"])
        worst = max(worst, r["worst"])
    assert worst < 1e-3, worst


def test_multiobj_graph_streams_and_step(c3):
    mt, frames, o, refs, (C, U, S, T) = c3
    fr = torch.from_numpy(frames).cuda()
    eager = [t.clone() for t in mt.track_windows(fr, graph=False)]
    g1 = [t.clone() for t in mt.track_windows(fr, graph=True)]
    g2 = [t.clone() for t in mt.track_windows(fr, graph=True)]          # replay
    for name, a, b, c in zip(("trk_logits", "boxes", "counts"), eager, g1, g2):
        if name == "boxes":                                         # rows past the count are scratch
            n = eager[2].cpu().numpy()
            for i in range(S * T):
                assert torch.equal(a[i, :n[i]], b[i, :n[i]]) and torch.equal(a[i, :n[i]], c[i, :n[i]]), (name, i)
        else:
            assert torch.equal(a, b), (name, (a.float() - b.float()).abs().max().item())
            assert torch.equal(a, c), (name, (a.float() - c.float()).abs().max().item())
    # streams are independent: stream 1 alone gives the same numbers as stream 1 next to stream 0
    alone = mt.track_windows(fr[1:2].contiguous(), graph=False)[0].clone()
    assert (alone - eager[0][T:]).abs().max().item() < 1e-3            # different K splits at batch 4 vs 8 (4.0e-4 on |logit| <= 12)
    # online stepping, two interleaved streams with persistent state == the windows
    mt.reset()
    for t in range(T):
        for s in (1, 0):
            bx = mt.step(frames[s, t], stream=s)
            n = int(eager[2][s * T + t].cpu())
            ref_rows = eager[1][s * T + t, :n].cpu().numpy()
            r = compare_rows(np.array([[b.x, b.y, b.w, b.h, b.c, b.get_score(), b.get_label(), -1] for b in bx]),
                             ref_rows, by_position=True)
            assert not r["unexplained"] and r["worst"] < 1e-3, (s, t, r)
    # a fifth step of stream 0 starts a new window (state reset every SEQUENCE_LENGTH steps)
    again = mt.step(frames[0, 0], stream=0)
    n0 = int(eager[2][0].cpu())
    assert abs(len(again) - n0) <= 2


def test_multiobj_plugin_predict_against_oracle(tmp_path):
    """MultiObjDetTracker.predict (image files in, annotated files out) against the oracle chain on the same
    resized frames -- not against itself."""
    import cv2
    from object_tracking_b200.models_tracking.MultiObjDetTracker import MultiObjDetTracker
    rng = np.random.default_rng(5)
    paths = []
    for i in range(4):
        p = str(tmp_path / f"f{i}.png")
        cv2.imwrite(p, rng.integers(0, 256, (300, 400, 3), dtype=np.uint8))
        paths.append(p)
    U = 64
    mt = MultiObjDetTracker(convlstm_units=U)
    C = mt.CLASS
    assert C == 12 and mt.detector.BATCH_SIZE == 4
    outs = [str(tmp_path / f"t{i}.png") for i in range(4)]
    trk = mt.predict(paths, outs)
    assert len(trk) == 4 and all(os.path.exists(p) for p in outs)
    x = np.stack([cv2.resize(cv2.imread(p), (416, 416)) for p in paths])
    wd = W.synthetic_yolo_weights(C, seed=0)
    wl = {k: v.astype(np.float64) for k, v in W.synthetic_multiobj_weights(C, U, seed=2).items()}
    o = yolo_oracle.yolo_forward(yolo_oracle.normalize(x), wd, C, dtype=np.float64)
    h = np.zeros((13, 13, U)); c = np.zeros((13, 13, U))
    total = 0
    for t in range(4):
        out, h, c = tracker_oracle.multiobj_step(o["logits"][t].reshape(13, 13, -1), o["feat"][t], h, c, wl)
        ref = decode_oracle.boxes_to_array(decode_oracle.decode_netout(
            out.reshape(13, 13, 5, 5 + C).astype(np.float32), 0.5, 0.45, W.ANCHORS, C))
        got = np.array([[b.x, b.y, b.w, b.h, b.c, b.get_score(), b.get_label(), -1] for b in trk[t]]).reshape(-1, 8)
        r = compare_rows(got, ref, by_position=True)
        assert not r["unexplained"], (t, r["unexplained"])
        assert r["worst"] < 1e-3 and r["matched"] >= len(ref) - 3
        total += r["matched"]
    assert total > 40


# ------------------------------------------------------------------------------------------------ C5
def test_c5_608_batch8_heatmap_tracker_pool_max():
    """BASELINE config 5 on one GPU: YOLOv2-608 C=80 + TinyHeatmapTracker, 8 streams, pool='Max'
    (MaxPooling2D(4,4)+Flatten: 19x19x1024 -> 16384 features, TinyTracker.py:29-33)."""
    from object_tracking_b200.models_tracking.TinyHeatmapTracker import TinyHeatmapTracker
    cfg = {k: dict(v) for k, v in TRACKER_CFG.items()}
    cfg["model_tracker"]["name"] = "TinyHeatmapTracker"
    cfg["model_tracker"]["sequence_length"] = 1
    cfg["train"]["pool"] = "Max"
    S, T = 8, 1
    trk = TinyHeatmapTracker(cfg, max_streams=S, detector_kwargs={"image_size": 608})
    eng = trk.model_detector.engine
    frames = np.random.default_rng(608).integers(0, 256, (S, T, 608, 608, 3), dtype=np.uint8)
    y = trk.track_windows(torch.from_numpy(frames).cuda(), graph=False).cpu().numpy()
    assert y.shape == (S, T, 1024)
    fv, _, heat, chosen = trk._decode_and_pool(S, 608, 608, 32)
    w = W.synthetic_detector_weights(80, seed=0)
    o = yolo_oracle.yolo_forward(yolo_oracle.normalize(frames.reshape(S, 608, 608, 3)), w, 80, dtype=np.float64,
                                 mode="darknet", want=["norm_20"])
    lg = eng.logits(S).cpu().numpy()
    assert lg.shape == (S, 19, 19, 5, 85)
    assert np.abs(lg - o["logits"]).max() < 1e-3
    n_feat = 4 * 4 * 1024
    wl = {k: v.astype(np.float64) for k, v in W.synthetic_lstm_weights(n_feat + 1024, 512, 1024, seed=1).items()}
    heat, fv = heat.cpu().numpy(), fv.cpu().numpy()
    assert fv.shape == (S, n_feat)
    for s in range(S):
        ref_fv = tracker_oracle.pool_features(o["norm_20"][s], "Max")
        assert np.abs(fv[s] - ref_fv).max() < 3e-3
        h = np.zeros((1, 512)); c = np.zeros((1, 512))
        yy, h, c = tracker_oracle.tracker_step(ref_fv[None], heat[s][None].astype(np.float64), h, c, wl)
        assert np.abs(y[s, 0] - yy[0]).max() < 1e-3, s
    assert (chosen.cpu().numpy() >= 0).sum() >= 2


# ------------------------------------------------------------------------------------------------ small-batch schedule
@pytest.mark.parametrize("B", [1, 3, 8])
def test_chain_schedule_matches_per_layer_schedule(B):
    """conv_chain_kernel (conv_2..23 in one persistent cooperative launch, grid barrier between layers, split-K finished
    in place) against the one-kernel-per-layer schedule and the fp64 oracle, every kept layer."""
    C = 2
    w = W.synthetic_yolo_weights(C, seed=0)
    frames = np.random.default_rng(40 + B).integers(0, 256, (B, 416, 416, 3), dtype=np.uint8)
    fr = torch.from_numpy(frames).cuda()
    a = _engine(n_class=C, max_batch=B, chain_max_batch=8)       # chain (library default: batch 1 only)
    b = _engine(n_class=C, max_batch=B, chain_max_batch=-1)      # never
    for e in (a, b):
        e.set_weights(w)
        e.finalize()
    la, lb = a.forward(fr).clone(), b.forward(fr).clone()
    assert a.launches < b.launches - 15                          # 3 launches instead of ~40
    assert (la - lb).abs().max().item() < 3e-4                   # conv_2 / conv_4 take a different kernel, same maths
    names = [f"norm_{i}" for i in (3, 4, 6, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20)] + ["concat", "conv_feat"]
    o = yolo_oracle.yolo_forward(yolo_oracle.normalize(frames), w, C, dtype=np.float64, want=names)
    rels = {}
    for n in names:
        got = a.extract(n, B).cpu().numpy()
        ref = o[n] if n != "conv_feat" else o["feat"]
        rels[n] = float(np.abs(got - ref).max() / np.abs(ref).max())
    assert max(rels.values()) < 4e-5, rels                      # 3.0e-5 at batch 8 (longer accumulation chains than at batch 2)
    assert np.abs(la.cpu().numpy() - o["logits"]).max() < 7e-4   # 5.0e-4 measured at batch 8 on |logit| <= 15
    # replays are bit-identical (fixed-order split-K finish, no atomics on data)
    assert torch.equal(a.forward(fr), la)


# ------------------------------------------------------------------------------------------------ C-ABI: graphs, broadcast
def test_c_abi_graph_capture_and_weight_broadcast():
    """b2t_graph_begin/end/launch (a plain-C host replays a whole step with one launch) and b2t_broadcast_weights
    (ncclBroadcast of the packed blob; exercised here on a one-rank communicator made with NCCL's own C API)."""
    import ctypes as C
    from object_tracking_b200 import _native as N
    B, Cc = 2, 2
    eng = _engine(n_class=Cc, max_batch=B)
    eng.set_weights(W.synthetic_yolo_weights(Cc, seed=0))
    eng.finalize()
    lib = eng.lib
    fr = torch.from_numpy(np.random.default_rng(3).integers(0, 256, (B, 416, 416, 3), dtype=np.uint8)).cuda()
    s = torch.cuda.Stream()
    anchors = (C.c_float * 10)(*W.ANCHORS)
    boxes = torch.zeros((B, 845, 8), device="cuda")
    counts = torch.zeros((B,), dtype=torch.int32, device="cuda")
    with torch.cuda.stream(s):
        ref_logits = eng.forward(fr).clone()
        rb, rc = eng.decode(eng.logits(B), 0.5, 0.45)
        rb, rc = rb.clone(), rc.clone()
        s.synchronize()
        g = C.c_void_p()
        n_before = lib.b2t_launch_count(eng.h)
        N.check(lib.b2t_graph_begin(eng.h, s.cuda_stream))
        N.check(lib.b2t_yolo_forward(eng.h, fr.data_ptr(), N.FRAME_U8, B, None, s.cuda_stream))
        N.check(lib.b2t_decode_nms(eng.h, eng.logits(B).data_ptr(), B, 13, 13, 5, Cc, 0.5, 0.45, anchors,
                                   boxes.data_ptr(), counts.data_ptr(), 845, s.cuda_stream))
        N.check(lib.b2t_graph_end(eng.h, s.cuda_stream, C.byref(g)))
        assert lib.b2t_launch_count(eng.h) == n_before            # nothing ran during the capture
        eng.logits(B).zero_()
        for _ in range(3):
            N.check(lib.b2t_graph_launch(g, s.cuda_stream))
        s.synchronize()
        per_replay = (lib.b2t_launch_count(eng.h) - n_before) // 3
        assert per_replay >= 20
        assert torch.equal(eng.logits(B), ref_logits) and torch.equal(counts, rc)
        for i in range(B):
            assert torch.equal(boxes[i, :int(rc[i])], rb[i, :int(rc[i])])
        lib.b2t_graph_destroy(g)
        assert lib.b2t_graph_begin(eng.h, None) < 0               # the legacy default stream cannot be captured
    nccl = C.CDLL("libnccl.so.2")
    comm = C.c_void_p()
    assert nccl.ncclCommInitAll(C.byref(comm), 1, (C.c_int * 1)(torch.cuda.current_device())) == 0
    before = eng.blob.clone()
    N.check(lib.b2t_broadcast_weights(eng.h, comm, 0, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(eng.blob, before)
    nccl.ncclCommDestroy(comm)
    assert lib.b2t_broadcast_weights(eng.h, None, 0, None) < 0


# ------------------------------------------------------------------------------------------------ formats (8f rank 2)
def test_tracker_checkpoints_and_voc_cfg(tmp_path):
    """Tracker heads from checkpoint files (BaseTracker.py:74-80 naming; .npz here, Keras .hdf5 through hdf5_lite) give
    the same numbers as the same weights passed as a dict; the compat layer loads a yolov2-voc.cfg-shaped network
    (20 classes, VOC anchors)."""
    import ctypes as C
    import copy
    from object_tracking_b200.models_tracking.TinyTracker import TinyTracker
    from object_tracking_b200.models_tracking.MultiObjDetTracker import MultiObjDetTracker
    cfg = copy.deepcopy(TRACKER_CFG)
    cfg["train"]["saved_model_dir"] = str(tmp_path) + "/"
    wl = W.synthetic_lstm_weights(1028, 512, 4, seed=21)
    W.save_tracker_checkpoint(str(tmp_path / "TinyTracker-CHKPNT-01-0.90.npz"), W.synthetic_lstm_weights(1028, 512, 4, seed=20))
    W.save_tracker_checkpoint(str(tmp_path / "TinyTracker-CHKPNT-05-0.40.npz"), wl)
    frames = torch.from_numpy(np.random.default_rng(2).integers(0, 256, (1, 4, 416, 416, 3), dtype=np.uint8)).cuda()
    a = TinyTracker(cfg, tracker_weights=wl, max_streams=1)
    ya = a.track_windows(frames, graph=False).clone()
    b = TinyTracker(cfg, max_streams=1)
    y0 = b.track_windows(frames, graph=False).clone()
    assert b.load_weights().endswith("CHKPNT-05-0.40.npz")
    yb = b.track_windows(frames, graph=False)
    assert torch.equal(ya, yb) and not torch.equal(y0, yb)
    del a, b
    # MultiObjDetTracker.load_weights: ConvLSTM + head + detector layers out of one checkpoint
    Cn, U = 2, 64
    wd, wt = W.synthetic_yolo_weights(Cn, seed=5), W.synthetic_multiobj_weights(Cn, U, seed=6)
    arrays = {f"/model_weights/tconv_lstm/tconv_lstm/{k}:0": wt[k] for k in ("kernel", "recurrent_kernel", "bias")}
    arrays["/model_weights/timedist_tconv2/timedist_tconv2/kernel:0"] = wt["head_kernel"]
    arrays["/model_weights/timedist_tconv2/timedist_tconv2/bias:0"] = wt["head_bias"]
    for s in W.yolo_layer_table(Cn):
        arrays[f"/model_weights/timedist_bbox/conv_{s.index}/kernel:0"] = wd[f"kernel_{s.index}"]
        if s.bn:
            for kn, ours in (("gamma", "gamma"), ("beta", "beta"), ("moving_mean", "mean"), ("moving_variance", "var")):
                arrays[f"/model_weights/timedist_bbox/norm_{s.index}/{kn}:0"] = wd[f"{ours}_{s.index}"]
        else:
            arrays[f"/model_weights/timedist_bbox/conv_{s.index}/bias:0"] = wd[f"bias_{s.index}"]
    ck = str(tmp_path / "MultiObjDetTracker-CHKPNT-03-0.55.npz")
    np.savez(ck, **arrays)
    fr4 = frames[0]
    m1 = MultiObjDetTracker({"LABELS": ["a", "b"]}, detector_weights=wd, tracker_weights=wt, convlstm_units=U)
    ref = m1.track_windows(fr4[None], graph=False)[0].clone()
    m2 = MultiObjDetTracker({"LABELS": ["a", "b"]}, convlstm_units=U)
    m2.load_weights(ck)
    assert m2.INITIAL_EPOCH == 3
    assert torch.equal(m2.track_windows(fr4[None], graph=False)[0], ref)
    del m1, m2
    # yolov2-voc.cfg: same graph, 20 classes, its own anchors
    import test_gpu_darknet_abi as T
    from oracle import darknet_ref
    d = str(tmp_path)
    cfgp, wts = os.path.join(d, "yolov2-voc.cfg"), os.path.join(d, "voc.weights")
    darknet_ref.write_yolov2_cfg(cfgp, 20, 416)
    txt = open(cfgp).read().replace("0.57273, 0.677385, 1.87446, 2.06253, 3.33843, 5.47434, 7.88282, 3.52778, 9.77052, 9.16828",
                                    "1.3221, 1.73145, 3.19275, 4.00944, 5.05587, 8.09892, 9.47112, 4.84053, 11.2364, 10.0071")
    assert "1.3221" in txt
    open(cfgp, "w").write(txt)
    W.write_darknet_weights(wts, W.synthetic_yolo_weights(20, seed=0), 20, major=0, minor=2, seen=12345)   # v0.2 header
    lib = T.bind(C.CDLL(os.path.join(os.path.dirname(W.__file__), "libb200track.so")))
    net = lib.load_network(cfgp.encode(), wts.encode(), 0)
    assert net and lib.network_width(net) == 416
    dd = lib.layer_dims(net, 31)
    assert (dd.h, dd.w, dd.c) == (13, 13, 125)
    if darknet_ref.available():
        ref_lib = T.bind(C.CDLL(darknet_ref.LIB_PATH))
        cwd = os.getcwd()
        net_r = ref_lib.load_network(cfgp.encode(), wts.encode(), 0)
        os.chdir(cwd)
        names = os.path.join(d, "voc.names")
        open(names, "w").write("\n".join(f"c{i}" for i in range(20)) + "\n")
        data = os.path.join(d, "voc.data")
        open(data, "w").write(f"classes= 20\nnames = {names}\n")
        meta = lib.get_metadata(data.encode())
        frame = np.random.default_rng(9).integers(0, 256, (416, 416, 3), dtype=np.uint8)
        ra = T.detect(ref_lib, net_r, meta, T.as_image(frame), thresh=.3)
        rb = T.detect(lib, net, meta, T.as_image(frame), thresh=.3)
        assert [x[0] for x in ra] == [x[0] for x in rb]
        for (n1, p1, b1), (n2, p2, b2) in zip(ra, rb):
            assert abs(p1 - p2) < 2e-4 and np.abs(np.array(b1) - np.array(b2)).max() < 5e-2


def test_strided_window_ingest_equals_copy_path():
    """Windows that are views of longer device-resident clips are gathered by the ingest kernel (b2t_ingest_frames), no
    staging copy: same numbers as the eager path on a contiguous copy, serial and pipelined, TinyTracker and
    MultiObjDetTracker."""
    from object_tracking_b200.models_tracking.TinyTracker import TinyTracker
    from object_tracking_b200.models_tracking.MultiObjDetTracker import MultiObjDetTracker
    clip = torch.from_numpy(np.random.default_rng(31).integers(0, 256, (2, 12, 416, 416, 3), dtype=np.uint8)).cuda()
    trk = TinyTracker(TRACKER_CFG, max_streams=2)
    eng = trk.model_detector.engine
    for j in (0, 4, 8):
        win = clip[:, j:j + 4]
        assert not win.is_contiguous() and eng.can_ingest(win)
        ref = trk.track_windows(win.contiguous(), graph=False).clone()
        assert torch.equal(trk.track_windows(win), ref)
        y = trk.track_windows(win, pipeline=True)
        with torch.cuda.stream(trk.tail_stream):
            yp = y.clone()
        torch.cuda.synchronize()
        assert torch.equal(yp, ref)
    assert not eng.can_ingest(clip[:, ::2][:, :4])              # frames of a stream must be consecutive in memory
    del trk
    mt = MultiObjDetTracker({"LABELS": ["a", "b"]}, convlstm_units=64, max_streams=2)
    win = clip[:, 3:7]
    ref = [t.clone() for t in mt.track_windows(win.contiguous(), graph=False)]
    got = mt.track_windows(win)
    assert torch.equal(got[0], ref[0]) and torch.equal(got[2], ref[2])


# ------------------------------------------------------------------------------------------------ callers after the path (8f rank 3)
def test_device_overlay_and_overlap_metric():
    """b2t_draw_boxes == cv2.rectangle(..., thickness 3) on decoded boxes (utils.draw_boxes without the text label),
    b2t_overlap_scores == utils.overlap_score / average_overlap_score in float64, bit for bit."""
    import cv2
    from oracle import overlay_oracle
    from object_tracking_b200.utility import utils as U
    eng = _engine(n_class=2, max_batch=1)
    rng = np.random.default_rng(12)
    B, H, W, M = 3, 300, 420, 40
    frames = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    rows = np.zeros((B, M, 8), np.float32)
    rows[..., 0:2] = rng.uniform(-0.05, 1.05, (B, M, 2))
    rows[..., 2:4] = rng.uniform(0.0, 0.6, (B, M, 2))
    rows[0, 0, :4] = (0.5, 0.5, 0.0, 0.0)                              # degenerate: a point
    rows[0, 1, :4] = (0.5, 0.5, 1.5, 1.5)                              # larger than the frame: nothing visible but clipping
    counts = np.array([M, 7, 0], np.int32)
    got = eng.draw_boxes(torch.from_numpy(frames.copy()).cuda(), torch.from_numpy(rows).cuda(),
                         torch.from_numpy(counts).cuda()).cpu().numpy()
    for b in range(B):
        ref = frames[b].copy()
        for r in rows[b, :counts[b]]:
            xa, ya, xb, yb = overlay_oracle.box_corners(r, W, H)
            cv2.rectangle(ref, (xa, ya), (xb, yb), (0, 255, 0), 3)
        assert np.array_equal(got[b], ref), b
    assert np.array_equal(got[2], frames[2])
    n = 257
    t = rng.uniform(0, 1, (n, 4)); p = rng.uniform(0, 1, (n, 4))
    t[:, 2:] += t[:, :2]; p[:, 2:] += p[:, :2]
    s, m = eng.overlap_scores(torch.from_numpy(t).cuda(), torch.from_numpy(p).cuda())
    ref_s = np.array([U.overlap_score(t[i], p[i]) for i in range(n)])
    assert np.array_equal(s.cpu().numpy(), ref_s)
    assert float(m.cpu()[0]) == U.average_overlap_score(t, p)


def test_decode_overflow_raises_everywhere():
    """the decode kernel reports a candidate-table overflow (thresholds far below 0.5: several classes per
    anchor pass) as count = -1; every host path must raise instead of slicing rows[:-1]."""
    from object_tracking_b200._native import B2TError
    from object_tracking_b200.engine import rows_to_host
    from object_tracking_b200.utility import utils as U
    eng = _engine(n_class=20, max_batch=1)
    net = np.zeros((13, 13, 5, 25), np.float32)
    net[..., 4] = 8.0                                             # conf ~ 1, uniform classes: p = 0.05 for 845 x 20 pairs
    boxes, counts = eng.decode(torch.from_numpy(net[None]).cuda(), 0.01, 0.45)
    assert int(counts.cpu()[0]) == -1
    with pytest.raises(B2TError):
        rows_to_host(boxes, counts)
    with pytest.raises(B2TError):
        U.decode_netout(net, 0.01, 0.45, W.ANCHORS, 20, engine=eng)
    ok = U.decode_netout(net, 0.5, 0.45, W.ANCHORS, 20, engine=eng)   # the reference's own threshold: nothing passes
    assert ok == []


def _wide_bn_weights(C, seed=77):
    """He-initialised kernels with BatchNorm statistics spread over orders of magnitude per channel: gamma log-uniform
    in [0.05, 20], variance in [0.01, 100] (folded scale 5e-3 .. 2e2 before normalisation); each layer's folded scales
    are then normalised to unit rms so that, like in a trained network, the activations stay O(1)..O(100) layer after
    layer while single channels sit three decades below or one above."""
    w = W.synthetic_yolo_weights(C, seed=3)
    rng = np.random.default_rng(seed)
    for s in W.yolo_layer_table(C):
        if s.bn:
            g = np.exp(rng.uniform(np.log(0.05), np.log(20.0), s.cout))
            v = np.exp(rng.uniform(np.log(0.01), np.log(100.0), s.cout))
            g /= np.sqrt(np.mean((g / np.sqrt(v + 1e-3)) ** 2))
            w[f"gamma_{s.index}"] = g.astype(np.float32)
            w[f"var_{s.index}"] = v.astype(np.float32)
            w[f"mean_{s.index}"] = (rng.standard_normal(s.cout) * 0.5).astype(np.float32)
    return w


def test_wide_batchnorm_spread_weights():
    """parity was only measured on He-initialised weights with tame BatchNorm statistics; trained YOLOv2
    weights have BN scales spread over orders of magnitude.  With _wide_bn_weights the fp16 (hi, lo) planes must keep
    their 22 bits through that spread: every checked layer within 5e-5 of the fp64 oracle relative to its own magnitude,
    logits within 1e-4 of theirs, darknet and Keras BatchNorm folding."""
    C, B = 2, 2
    w = _wide_bn_weights(C)
    frames = np.random.default_rng(5).integers(0, 256, (B, 416, 416, 3), dtype=np.uint8)
    names = [f"norm_{i}" for i in (3, 6, 9, 13, 14, 18, 20)] + ["concat"]
    for sem in ("keras", "darknet"):
        o = yolo_oracle.yolo_forward(yolo_oracle.normalize(frames), w, C, dtype=np.float64, want=names, mode=sem)
        eng = _engine(n_class=C, max_batch=B, semantics=sem)
        eng.set_weights(w)
        eng.finalize()
        lg = eng.forward(torch.from_numpy(frames).cuda()).cpu().numpy()
        stats = {}
        for n in names + ["conv_feat"]:
            got = eng.extract(n, B).cpu().numpy()
            ref = o[n] if n != "conv_feat" else o["feat"]
            assert np.isfinite(got).all(), (sem, n)
            ch = np.abs(ref).reshape(-1, ref.shape[-1]).max(axis=0)
            stats[n] = (float(np.abs(ref).max()), float(np.abs(got - ref).max() / np.abs(ref).max()),
                        float(ch.max() / max(ch[ch > 0].min(), 1e-30)))
        assert max(v[1] for v in stats.values()) < 5e-5, (sem, stats)
        assert max(v[0] for v in stats.values()) < 6e4, (sem, stats)                 # inside the fp16 planes' range
        assert max(v[2] for v in stats.values()) > 100, (sem, stats)                  # the per-channel spread is real
        err = np.abs(lg - o["logits"]).max()
        assert err < 1e-4 * max(1.0, np.abs(o["logits"]).max()), (sem, err, np.abs(o["logits"]).max())
        del eng


@pytest.mark.parametrize("size,B", [(320, 3), (352, 1), (512, 2), (544, 1)])
def test_other_network_sizes_against_fp64_oracle(size, B):
    """darknet accepts any width/height that is a multiple of 32 (cfg `width=`/`height=`, parser.c:611-623; its multi-scale
    training uses 320..608): grids 10, 11, 16 and 17 exercise other tile geometries (odd grids, partial tiles) and, at one
    frame, the chain schedule's image-size condition.  Logits against the fp64 oracle, every kept layer of frame 0."""
    from object_tracking_b200.engine import DetectorEngine
    w = W.synthetic_yolo_weights(20, seed=4)
    frames = np.random.default_rng(size).integers(0, 256, (B, size, size, 3), dtype=np.uint8)
    e = DetectorEngine(n_class=20, image_size=size, max_batch=B, semantics="darknet")
    e.set_weights(w)
    e.finalize()
    lg = e.forward(torch.from_numpy(frames).cuda()).cpu().numpy()
    G = size // 32
    assert lg.shape == (B, G, G, 5, 25)
    names = ["norm_3", "norm_7", "norm_13", "norm_20", "concat"]
    o = yolo_oracle.yolo_forward(yolo_oracle.normalize(frames), w, 20, dtype=np.float64, mode="darknet", want=names)
    assert np.abs(lg - o["logits"]).max() < 1e-3 * max(1.0, np.abs(o["logits"]).max() / 10), np.abs(lg - o["logits"]).max()
    for n in names:
        got, ref = e.extract(n, B).cpu().numpy(), o[n]
        assert got.shape == ref.shape, n
        assert np.abs(got - ref).max() / np.abs(ref).max() < 3e-5, (n, size)
    dets, counts = e.region_detect(torch.from_numpy(lg).cuda(), 0.3, 0.45, size, size)
    region = darknet_oracle.region_forward(np.transpose(o["logits"][0].reshape(G, G, -1), (2, 0, 1)).astype(np.float32), 20)
    boxes, obj, prob = darknet_oracle.detect(region, size, size, size, size, 0.3, 0.45, 20)
    kept = int((prob > 0).sum())                                                    # rows = (box, class) pairs above the threshold
    assert abs(int(counts.cpu()[0]) - kept) <= 1, (int(counts.cpu()[0]), kept)     # a score on the threshold may flip
