"""CPU suite, part 1: the oracle against the committed golden vectors (which hold outputs of the
REFERENCE's own code: utility/utils.py decode_netout exec'd from /root/reference, and the reference's
darknet C library), plus host-side weight IO."""
import os
import tempfile

import numpy as np
import pytest

from oracle import darknet_oracle, decode_oracle, tracker_oracle, yolo_oracle
from oracle.cases import DECODE_KINDS, decode_case
from object_tracking_b200 import weights as W

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_decode_oracle_matches_reference_outputs():
    z = np.load(os.path.join(GOLD, "decode_cases.npz"))
    n = int(z["n_cases"])
    assert n >= 80
    nonempty = 0
    for i in range(n):
        g, c, kind, seed = (int(v) for v in z[f"meta_{i}"])
        net = decode_case(seed, g, c, DECODE_KINDS[kind])
        assert np.isclose(net.astype(np.float64).sum(), float(z[f"insum_{i}"]), rtol=0, atol=1e-6)
        got = decode_oracle.boxes_to_array(decode_oracle.decode_netout(net, 0.5, 0.45, W.ANCHORS, c))
        ref = z[f"box_{i}"]
        assert got.shape == ref.shape, (i, kind)
        assert np.array_equal(got, ref), (i, kind)          # bit-exact incl. order, labels, anchor ids
        nonempty += len(ref) > 0
    assert nonempty > n // 2


def test_decode_edge_cases():
    # empty, a single confident anchor, two identical boxes of one class (IoU = 1 >= thr -> one survives)
    net = np.full((13, 13, 5, 7), -20.0, np.float32)
    assert decode_oracle.decode_netout(net, 0.5, 0.45, W.ANCHORS, 2) == []
    net[3, 4, 1, 4] = 8.0; net[3, 4, 1, 5] = 9.0; net[3, 4, 1, 0:4] = 0.0
    out = decode_oracle.decode_netout(net, 0.5, 0.45, W.ANCHORS, 2)
    assert len(out) == 1 and out[0].get_label() == 0 and out[0].cell == (3 * 13 + 4) * 5 + 1
    net[3, 4, 2] = net[3, 4, 1]
    net[3, 4, 2, 2:4] += np.log(np.float32([W.ANCHORS[2] / W.ANCHORS[4], W.ANCHORS[3] / W.ANCHORS[5]]))
    out = decode_oracle.decode_netout(net, 0.5, 0.45, W.ANCHORS, 2)
    assert len(out) == 1


def test_darknet_oracle_matches_reference_library_outputs():
    z = np.load(os.path.join(GOLD, "darknet_416.npz"))
    region = darknet_oracle.region_forward(z["logits"], 80)
    assert np.abs(region - z["region"]).max() < 1e-5
    boxes, obj, prob = darknet_oracle.detect(z["region"], 416, 416, 416, 416, 0.5, 0.45, 80)
    live = np.nonzero(obj)[0]
    assert live.size == z["det_obj"].shape[0]
    assert np.allclose(boxes[live], z["det_boxes"], atol=1e-3)
    assert np.allclose(obj[live], z["det_obj"], atol=1e-6)
    assert np.allclose(prob[live], z["det_prob"], atol=1e-6)
    assert float(z["oracle_err"].max()) < 2e-3          # fp64 forward restatement vs libdarknet, recorded at pin time


def test_yolo_oracle_regression_vector():
    z = np.load(os.path.join(GOLD, "keras_416_c2.npz"))
    w = W.synthetic_yolo_weights(2, seed=int(z["weight_seed"]))
    frames = np.random.default_rng(int(z["frame_seed"])).integers(0, 256, (2, 416, 416, 3), dtype=np.uint8)
    o = yolo_oracle.yolo_forward(yolo_oracle.normalize(frames[:1]).astype(np.float32), w, 2, dtype=np.float32)
    assert np.abs(o["logits"][0] - z["logits"][0]).max() < 2e-3       # fp32 run vs committed fp64 vector
    assert np.abs(o["feat"][0].max(axis=(0, 1)) - z["feat_globalmax"][0]).max() < 5e-3


def test_space_to_depth_orderings_differ():
    x = np.arange(64 * 26 * 26, dtype=np.float64).reshape(1, 64, 26, 26)
    import torch
    t = torch.from_numpy(x)
    tf_order = yolo_oracle.space_to_depth_x2(t.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
    dk = yolo_oracle.darknet_reorg(t, 2)
    assert tf_order.shape == dk.shape == (1, 256, 13, 13)
    assert not torch.equal(tf_order, dk)
    assert torch.equal(torch.sort(tf_order.reshape(-1))[0], torch.sort(dk.reshape(-1))[0])   # both permutations


def test_tracker_oracle_regression_vectors():
    z = np.load(os.path.join(GOLD, "tracker_cases.npz"))
    out = tracker_oracle.make_cases()
    for k in z.files:
        assert np.allclose(out[k], z[k], rtol=0, atol=1e-12), k


def test_heatmap_helpers():
    h = tracker_oracle.generate_heatmap_feat(0.25, 0.5, 0.25, 0.125, 32).reshape(32, 32)
    assert h.sum() == (4 + 1) * (8 + 1) and h[16, 8] == 1 and h[15, 8] == 0
    assert tracker_oracle.generate_rectangle_from_heatmap(h) == (8, 16, 16, 20)
    assert tracker_oracle.generate_rectangle_from_heatmap(np.zeros((32, 32))) == (32, 32, -1, -1)


def test_darknet_weight_file_roundtrip():
    w = W.synthetic_yolo_weights(2, seed=3)
    with tempfile.TemporaryDirectory() as d:
        for (major, minor) in ((0, 1), (0, 2)):          # int32 vs size_t "seen" header (parser.c:1220-1226)
            p = os.path.join(d, f"v{major}{minor}.weights")
            W.write_darknet_weights(p, w, 2, major=major, minor=minor)
            r = W.read_darknet_weights(p, 2)
            assert set(r) == set(w)
            for k in w:
                assert np.array_equal(r[k], w[k]), k
        with pytest.raises(ValueError):
            W.read_darknet_weights(p, 80)                  # wrong class count -> truncated
    assert W.n_params(80) == 50_983_561                  # 50 942 217 + 4 * sum(Cout of the 22 BN layers)


def test_param_and_traffic_counts_match_survey():
    # W = 50 942 217 conv weights + conv_23 bias (BN vectors excluded), R, Wr
    t = W.traffic_model(80, 416)
    assert t["W"] == 50_942_217
    assert t["R"] == 8_955_648
    assert t["Wr"] == 8_508_305
    assert abs(t["flops"] - 29.464e9) < 0.01e9
    assert abs(W.forward_bytes(80, 416, batch=1) - 273.6e6) < 0.1e6


def test_u8_normalisation_needs_no_table():
    # conv_1 computes float(u)/255.f in fp32; the reference computes image/255. in float64 and Keras casts to fp32
    u = np.arange(256)
    assert np.array_equal((u.astype(np.float64) / 255.0).astype(np.float32), u.astype(np.float32) / np.float32(255.0))


def test_ingest_oracle_matches_cv2_resize_goldens():
    """Frame ingest (KerasYOLO.py:526 cv2.resize): the restatement against outputs of OpenCV itself (committed), and
    against the installed cv2 when there is one."""
    from oracle import ingest_oracle
    z = np.load(os.path.join(GOLD, "resize_cases.npz"))
    for seed, h, w, dst in ingest_oracle.RESIZE_CASES:
        img = ingest_oracle.resize_case(seed, h, w)
        got = ingest_oracle.resize_linear_u8(img, dst, dst)
        assert np.array_equal(got, z[f"case{seed}"]), (seed, h, w, dst)
    try:
        import cv2
    except ImportError:
        return
    img = np.random.default_rng(3).integers(0, 256, (211, 333, 3), dtype=np.uint8)
    for dst in (416, 608):
        assert np.array_equal(ingest_oracle.resize_linear_u8(img, dst, dst), cv2.resize(img, (dst, dst)))


def test_heatmap_oracle_matches_reference_outputs():
    """Committed outputs of the REFERENCE's generate_heatmap_feat / generate_rectangle_from_heatmap
    (utility/utils.py:53-79 exec'd from /root/reference by oracle/make_golden.py), not of our restatement."""
    from oracle.cases import heatmap_case_inputs
    z = np.load(os.path.join(GOLD, "heatmap_cases.npz"))
    n = int(z["n"])
    xywh, heat = heatmap_case_inputs(n, int(z["seed"]))
    feats = np.unpackbits(z["feat_bits"], axis=1)[:, :1024]
    for i in range(n):
        got = tracker_oracle.generate_heatmap_feat(*xywh[i], hmap_size=32)
        assert np.array_equal(got, feats[i].astype(np.float64)), i
        assert tracker_oracle.generate_rectangle_from_heatmap(heat[i], 0.75, 32) == tuple(int(v) for v in z["rect"][i]), i
        assert tracker_oracle.generate_rectangle_from_heatmap(feats[i], 0.75, 32) == tuple(int(v) for v in z["rect_of_feat"][i]), i
    assert feats.sum() > 1000 and (z["rect"][:, 2] >= 0).sum() > n // 2


def test_lstm_oracle_matches_torch_lstmcell():
    """Independent pin of the LSTM restatement's gate order (i, f, c, o), weight layout (kernel (n_in,4u) =
    weight_ih^T, recurrent_kernel (u,4u) = weight_hh^T) and state update: torch.nn.LSTMCell computes the same
    cell with sigmoid gates (Keras >= 2.3's recurrent_activation); the hard_sigmoid of Keras 2.0-2.2 differs only
    in the gate non-linearity, checked separately against its definition."""
    import torch
    n_in, u, S = 37, 16, 5
    w = W.synthetic_lstm_weights(n_in, u, 4, seed=9)
    cell = torch.nn.LSTMCell(n_in, u).double()
    with torch.no_grad():
        cell.weight_ih.copy_(torch.from_numpy(w["kernel"].T.astype(np.float64)))
        cell.weight_hh.copy_(torch.from_numpy(w["recurrent_kernel"].T.astype(np.float64)))
        cell.bias_ih.copy_(torch.from_numpy(w["bias"].astype(np.float64)))
        cell.bias_hh.zero_()
    rng = np.random.default_rng(4)
    h = np.zeros((S, u)); c = np.zeros((S, u))
    ht, ct = torch.zeros(S, u, dtype=torch.float64), torch.zeros(S, u, dtype=torch.float64)
    w64 = {k: v.astype(np.float64) for k, v in w.items()}
    for _ in range(6):
        x = rng.standard_normal((S, n_in))
        h, c = tracker_oracle.lstm_step(x, h, c, w64, recurrent_activation=tracker_oracle.sigmoid)
        with torch.no_grad():
            ht, ct = cell(torch.from_numpy(x), (ht, ct))
        assert np.abs(h - ht.numpy()).max() < 1e-12 and np.abs(c - ct.numpy()).max() < 1e-12
    x = np.linspace(-4, 4, 33)
    assert np.array_equal(tracker_oracle.hard_sigmoid(x), np.clip(0.2 * x + 0.5, 0, 1))
    assert tracker_oracle.hard_sigmoid(np.array([-2.5, 0.0, 2.5])).tolist() == [0.0, 0.5, 1.0]


def test_convlstm_oracle_matches_torch_reference():
    """ConvLSTM2D restatement against an independent torch composition (conv2d with 'same' padding on NCHW
    tensors, gate slices i,f,c,o) -- checks the HWIO->OIHW kernel handling and the state update."""
    import torch
    import torch.nn.functional as F
    G, cin, u = 5, 7, 4
    w = W.synthetic_convlstm_weights(cin, u, 6, seed=5)
    rng = np.random.default_rng(1)
    h = np.zeros((G, G, u)); c = np.zeros((G, G, u))
    ht = torch.zeros(1, u, G, G, dtype=torch.float64); ct = torch.zeros(1, u, G, G, dtype=torch.float64)
    k = torch.from_numpy(w["kernel"].astype(np.float64)).permute(3, 2, 0, 1)
    r = torch.from_numpy(w["recurrent_kernel"].astype(np.float64)).permute(3, 2, 0, 1)
    b = torch.from_numpy(w["bias"].astype(np.float64))
    w64 = {kk: v.astype(np.float64) for kk, v in w.items()}
    for _ in range(3):
        z = rng.standard_normal((G, G, cin))
        h, c = tracker_oracle.convlstm_step(z, h, c, w64)
        zt = torch.from_numpy(z).permute(2, 0, 1)[None]
        g = F.conv2d(zt, k, b, padding=1) + F.conv2d(ht, r, padding=1)
        hs = lambda t: torch.clamp(0.2 * t + 0.5, 0, 1)
        i, f, o = hs(g[:, :u]), hs(g[:, u:2 * u]), hs(g[:, 3 * u:])
        ct = f * ct + i * torch.tanh(g[:, 2 * u:3 * u])
        ht = o * torch.tanh(ct)
        assert np.abs(h - ht[0].permute(1, 2, 0).numpy()).max() < 1e-12


def test_detection_to_tracker_input_spec():
    """preprocessing.py:434-456: the first (highest-probability) detection of the class-filtered list, normalised by
    the frame size; zeros when the list is empty; heat-map variant takes the top-left corner."""
    lst = [("car", 0.9, (208.0, 104.0, 41.6, 83.2)), ("person", 0.8, (10.0, 10.0, 5.0, 5.0))]
    v = tracker_oracle.detection_to_tracker_input(lst, 416, 208)
    assert np.allclose(v, [0.5, 0.5, 0.1, 0.4]) and v.dtype == np.float32
    assert np.array_equal(tracker_oracle.detection_to_tracker_input([], 416, 208), np.zeros(4, np.float32))
    hm = tracker_oracle.detection_to_tracker_input(lst, 416, 208, heatmap_size=32).reshape(32, 32)
    assert hm.sum() == (int(0.1 * 32) + 1) * (int(0.4 * 32) + 1) and hm[int(0.3 * 32), int(0.45 * 32)] == 1
    assert tracker_oracle.detection_to_tracker_input([], 416, 208, heatmap_size=32).sum() == 1    # the (0,0) cell


def test_planted_synthetic_detector_emits_allowed_classes():
    """weights.synthetic_detector_weights: the tracker's allowed classes (person, car) are planted in the random head
    so that the positive branch of the detection choice is exercised (it never was)."""
    a, b = W.synthetic_yolo_weights(80, seed=0), W.synthetic_detector_weights(80, seed=0)
    for k in a:
        if k != "bias_23":
            assert np.array_equal(a[k], b[k]), k
    d = (b["bias_23"] - a["bias_23"]).reshape(5, 85)
    assert np.allclose(d[:, 5], 10.0) and np.allclose(d[:, 7], 9.0) and np.count_nonzero(d) == 10


def test_overlay_oracle_matches_cv2_rectangle():
    """draw_boxes' cv2.rectangle(..., thickness 3) restated as a pixel-set rule == the installed OpenCV, incl. clipped,
    reversed and degenerate rectangles."""
    import cv2
    from oracle import overlay_oracle
    rng = np.random.default_rng(0)
    for _ in range(150):
        H, W = int(rng.integers(20, 90)), int(rng.integers(20, 90))
        a, b = np.zeros((H, W, 3), np.uint8), np.zeros((H, W, 3), np.uint8)
        for _ in range(3):
            x1, x2 = (int(v) for v in rng.integers(-15, W + 15, 2))
            y1, y2 = (int(v) for v in rng.integers(-15, H + 15, 2))
            cv2.rectangle(a, (x1, y1), (x2, y2), (0, 255, 0), 3)
            overlay_oracle.rectangle3(b, x1, y1, x2, y2, (0, 255, 0))
        assert np.array_equal(a, b)
