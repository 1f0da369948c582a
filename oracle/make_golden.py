"""Pin the oracle against the reference and write tests/golden/*.npz.
Runs ONLY in the build container (needs /root/reference and oracle/_ref/libdarknet.so).
TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden.py            # all
    python oracle/make_golden.py decode     # only the decode/NMS goldens

What is produced (all small, committed):
  decode_cases.npz     inputs + the outputs of the reference's OWN decode_netout
                       (utility/utils.py:113-270 exec'd from /root/reference, numpy 2.x)
  darknet_416.npz      outputs of the reference's darknet C library on the seeded synthetic
                       .weights + a seeded 416x416 frame: region-input logits (layer 31),
                       fv_layer 25 feature (max-pooled), post-NMS detections
  keras_416_c2.npz     fp64 oracle outputs for the Keras-semantics graph (C=2; regression vector)
  tracker_cases.npz    fp64 oracle outputs of the LSTM / heat-map / ConvLSTM steps
  jpeg/*.jpg, jpeg_cases.npz   synthetic JPEG files + what the reference library's load_image_color decodes from them
  heatmap_cases.npz    outputs of the reference's OWN generate_heatmap_feat / generate_rectangle_from_heatmap
                       (utility/utils.py:53-79 exec'd from /root/reference) on seeded boxes / heat-maps
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import decode_oracle, yolo_oracle, tracker_oracle, darknet_ref  # noqa: E402
from oracle.cases import decode_case, heatmap_case_inputs, DECODE_KINDS, DECODE_SPECS  # noqa: E402

pkg = importlib.import_module("object-tracking_b200")
W = importlib.import_module("object-tracking_b200.weights")

ANCHORS = W.ANCHORS


def load_reference_decode():
    """exec utility/utils.py:113-270 (BoundBox ... softmax); no py2-only syntax on those lines."""
    src = open(os.path.join(REF, "utility", "utils.py")).read().split("\n")
    ns = {"np": np}
    exec(compile("\n".join(src[112:270]), "reference:utility/utils.py", "exec"), ns)
    return ns["decode_netout"]


def make_decode_goldens():
    """Every fixture input is regenerated from its seed by oracle/cases.py::decode_case, so the
    npz only carries the REFERENCE outputs plus an input checksum."""
    ref_decode = load_reference_decode()
    cases, n_checked, idx = {}, 0, 0
    for rep in range(40):
        for (g, c) in DECODE_SPECS:
            for kind in DECODE_KINDS:
                if g == 19 and rep > 5:
                    continue
                seed = 900000 + 1000 * rep + 37 * DECODE_SPECS.index((g, c)) + DECODE_KINDS.index(kind)
                net = decode_case(seed, g, c, kind)
                ref = ref_decode(net.copy(), 0.5, 0.45, ANCHORS, c)
                mine = decode_oracle.decode_netout(net, 0.5, 0.45, ANCHORS, c)
                assert len(ref) == len(mine), (kind, g, c, len(ref), len(mine))
                for a, b in zip(ref, mine):
                    for f in "xywhc":
                        assert np.float32(getattr(a, f)) == np.float32(getattr(b, f)), (kind, f)
                    assert np.array_equal(np.asarray(a.classes), np.asarray(b.classes))
                    assert a.get_label() == b.get_label()
                n_checked += 1
                if rep < 3:                      # commit the reference outputs of a few of each
                    rows = np.zeros((len(ref), 8), np.float64)
                    for i, (a, b) in enumerate(zip(ref, mine)):
                        rows[i] = (a.x, a.y, a.w, a.h, a.c, a.get_score(), a.get_label(), b.cell)
                    cls = np.stack([np.asarray(a.classes) for a in ref]) if ref else np.zeros((0, c), np.float32)
                    cases[f"box_{idx}"] = rows
                    cases[f"cls_{idx}"] = cls.astype(np.float32)
                    cases[f"meta_{idx}"] = np.array([g, c, DECODE_KINDS.index(kind), seed])
                    cases[f"insum_{idx}"] = np.array(net.astype(np.float64).sum())
                    idx += 1
    cases["n_cases"] = np.array(idx)
    np.savez_compressed(os.path.join(GOLD, "decode_cases.npz"), **cases)
    print(f"[decode] oracle == reference decode_netout on {n_checked} tensors; {idx} fixtures written")


def make_darknet_golden():
    n_class = 80
    w = W.synthetic_yolo_weights(n_class, seed=0)
    tmp = tempfile.mkdtemp()
    wpath = os.path.join(tmp, "synthetic.weights")
    W.write_darknet_weights(wpath, w, n_class)
    # (a) the reference's own cfg, (b) our generated cfg -> must give identical outputs
    ref = darknet_ref.DarknetRef(os.path.join(REF, "darknet", "cfg", "yolov2.cfg"), wpath)
    cfg2 = os.path.join(tmp, "gen.cfg")
    darknet_ref.write_yolov2_cfg(cfg2, n_class, 416)
    ref2 = darknet_ref.DarknetRef(cfg2, wpath)
    rng = np.random.default_rng(1234)
    frame = rng.integers(0, 256, (416, 416, 3), dtype=np.uint8)
    chw = np.ascontiguousarray(np.transpose(frame.astype(np.float32) / np.float32(255.), (2, 0, 1)))
    ref.predict(chw)
    ref2.predict(chw)
    logits = ref.extract(31).reshape(425, 13, 13)          # layer 30 output = region input
    assert np.array_equal(logits, ref2.extract(31).reshape(425, 13, 13)), "generated cfg != yolov2.cfg"
    feat = ref.extract(25).reshape(1024, 13, 13)
    skip = ref.extract(27).reshape(64, 26, 26)             # conv_21 output (layer 26)
    reorg = ref.extract(28).reshape(256, 13, 13)
    region = ref.extract(32).reshape(425, 13, 13)
    assert ref.layer_dims(25) == (13, 13, 1024) and ref.layer_dims(26) == (26, 26, 512)
    boxes, obj, prob = ref.detect(416, 416, 0.5, 0.5, 0.45, n_class)

    # pin the restatement: fp64 oracle (darknet semantics) vs the library
    o = yolo_oracle.yolo_forward(np.transpose(chw, (1, 2, 0))[None].astype(np.float64), w, n_class,
                                 dtype=np.float64, mode="darknet", want=["norm_20", "norm_21", "concat"])
    lo = np.transpose(o["logits"].reshape(13, 13, 425), (2, 0, 1))
    fo = np.transpose(o["norm_20"][0], (2, 0, 1))      # fv_layer 25 -> layers[24] = conv_20
    so = np.transpose(o["norm_21"][0], (2, 0, 1))
    ro = np.transpose(o["concat"][0], (2, 0, 1))[:256]
    e_l, e_f = np.abs(lo - logits).max(), np.abs(fo - feat).max()
    e_s, e_r = np.abs(so - skip).max(), np.abs(ro - reorg).max()
    print(f"[darknet] |oracle64 - libdarknet| logits {e_l:.3e} (max |logit| {np.abs(logits).max():.2f}) "
          f"feat {e_f:.3e} conv_21 {e_s:.3e} reorg {e_r:.3e}")
    assert e_l < 2e-3 and e_f < 2e-3 and e_s < 1e-3 and e_r < 1e-3
    from oracle import darknet_oracle
    reg_o = darknet_oracle.region_forward(logits, n_class)
    print(f"[darknet] region layer restatement err {np.abs(reg_o - region).max():.3e}")
    assert np.abs(reg_o - region).max() < 1e-5
    b_o, obj_o, prob_o = darknet_oracle.detect(region, 416, 416, 416, 416, 0.5, 0.45, n_class)
    live_r = np.nonzero(obj)[0]
    live_o = np.nonzero(obj_o)[0]
    print(f"[darknet] detections after NMS: lib {live_r.size} oracle {live_o.size}")
    assert live_r.size == live_o.size
    assert np.allclose(boxes[live_r], b_o[live_o], atol=1e-3) and np.allclose(prob[live_r], prob_o[live_o], atol=1e-6)
    np.savez_compressed(
        os.path.join(GOLD, "darknet_416.npz"),
        frame_seed=np.array(1234), weight_seed=np.array(0),
        logits=logits.astype(np.float32), region=region.astype(np.float32),
        feat_globalmax=feat.reshape(1024, -1).max(1).astype(np.float32),
        feat_sub=feat[::16].astype(np.float32),
        reorg_sub=reorg[::8].astype(np.float32),
        det_boxes=boxes[live_r], det_obj=obj[live_r], det_prob=prob[live_r],
        oracle_err=np.array([e_l, e_f, e_s, e_r]))
    print("[darknet] golden written")


def make_keras_golden():
    n_class = 2
    w = W.synthetic_yolo_weights(n_class, seed=0)
    rng = np.random.default_rng(1234)
    frames = rng.integers(0, 256, (2, 416, 416, 3), dtype=np.uint8)
    x = yolo_oracle.normalize(frames)
    o64 = yolo_oracle.yolo_forward(x, w, n_class, dtype=np.float64)
    o32 = yolo_oracle.yolo_forward(x.astype(np.float32), w, n_class, dtype=np.float32)
    print(f"[keras] fp32-vs-fp64 noise floor: logits {np.abs(o64['logits'] - o32['logits']).max():.3e} "
          f"feat {np.abs(o64['feat'] - o32['feat']).max():.3e}; max|logit| {np.abs(o64['logits']).max():.2f}")
    np.savez_compressed(os.path.join(GOLD, "keras_416_c2.npz"),
                        frame_seed=np.array(1234), weight_seed=np.array(0),
                        logits=o64["logits"].astype(np.float32),
                        feat_globalmax=o64["feat"].max(axis=(1, 2)).astype(np.float32),
                        feat_sub=o64["feat"][:, :, :, ::32].astype(np.float32))
    print("[keras] golden written")


def make_tracker_golden():
    out = tracker_oracle.make_cases()
    np.savez_compressed(os.path.join(GOLD, "tracker_cases.npz"), **out)
    print(f"[tracker] {len(out)} arrays written")


def load_reference_heatmap():
    """exec utility/utils.py:53-79 (generate_heatmap_feat, generate_rectangle_from_heatmap): plain numpy, no
    py2-only syntax on those lines."""
    src = open(os.path.join(REF, "utility", "utils.py")).read().split("\n")
    ns = {"np": np}
    exec(compile("\n".join(src[52:79]), "reference:utility/utils.py", "exec"), ns)
    return ns["generate_heatmap_feat"], ns["generate_rectangle_from_heatmap"]


def make_heatmap_golden():
    ref_feat, ref_rect = load_reference_heatmap()
    bad = 0
    xywh, heat = heatmap_case_inputs(3000, 20250)
    for i in range(3000):
        a = ref_feat(*xywh[i], hmap_size=32)
        b = tracker_oracle.generate_heatmap_feat(*xywh[i], hmap_size=32)
        bad += int(not np.array_equal(a, b))
        ra = tuple(int(v) for v in ref_rect(heat[i], 0.75, 32))
        rb = tuple(int(v) for v in tracker_oracle.generate_rectangle_from_heatmap(heat[i], 0.75, 32))
        bad += int(ra != rb)
    assert bad == 0, f"heat-map oracle differs from the reference in {bad} cases"
    n = 96
    xywh, heat = heatmap_case_inputs(n, 777)
    feats = np.stack([ref_feat(*xywh[i], hmap_size=32) for i in range(n)])
    rects = np.array([ref_rect(heat[i], 0.75, 32) for i in range(n)], dtype=np.int32)
    rects_of_feats = np.array([ref_rect(feats[i].reshape(32, 32), 0.75, 32) for i in range(n)], dtype=np.int32)
    np.savez_compressed(os.path.join(GOLD, "heatmap_cases.npz"), seed=np.array(777), n=np.array(n),
                        feat_bits=np.packbits(feats.astype(np.uint8), axis=1), rect=rects, rect_of_feat=rects_of_feats)
    print(f"[heatmap] oracle == reference utils.py:53-79 on 3000 cases; {n} reference outputs written")


def make_resize_golden():
    """Outputs of the installed OpenCV's cv2.resize (the reference's ingest step, KerasYOLO.py:526) on seeded images."""
    import cv2
    from oracle import ingest_oracle
    out, bad = {"cv2_version": np.array(cv2.__version__)}, 0
    for seed, h, w, dst in ingest_oracle.RESIZE_CASES:
        img = ingest_oracle.resize_case(seed, h, w)
        ref = cv2.resize(img, (dst, dst))
        out[f"case{seed}"] = ref
        bad += int((ingest_oracle.resize_linear_u8(img, dst, dst) != ref).sum())
    # wider sweep, not stored: the restatement must reproduce cv2 bit for bit
    rng = np.random.default_rng(7)
    for h, w in [(576, 768), (480, 640), (1080, 1920), (417, 415), (100, 1000), (720, 1280), (300, 500)]:
        for dst in (416, 608):
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            bad += int((ingest_oracle.resize_linear_u8(img, dst, dst) != cv2.resize(img, (dst, dst))).sum())
    assert bad == 0, f"ingest oracle differs from cv2.resize in {bad} bytes"
    np.savez_compressed(os.path.join(GOLD, "resize_cases.npz"), **out)
    print(f"[resize] {len(out) - 1} cv2 outputs written (cv2 {cv2.__version__}); restatement bit-exact on the sweep")


def make_jpeg_golden():
    """JPEG fixtures (synthetic images written by OpenCV's libjpeg: every chroma subsampling, restart intervals, grey,
    optimised Huffman tables, tiny sizes) and what the REFERENCE library's load_image_color (stb_image path,
    image.c:1442-1482) decodes from them, as bytes (the library returns byte / 255.)."""
    import ctypes as C
    import cv2

    class IMAGE(C.Structure):
        _fields_ = [("w", C.c_int), ("h", C.c_int), ("c", C.c_int), ("data", C.POINTER(C.c_float))]

    ref = C.CDLL(darknet_ref.LIB_PATH)
    ref.load_image_color.restype = IMAGE
    ref.load_image_color.argtypes = [C.c_char_p, C.c_int, C.c_int]
    ref.free_image.argtypes = [IMAGE]
    d = os.path.join(GOLD, "jpeg")
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(2025)
    smooth = cv2.GaussianBlur(rng.integers(0, 256, (90, 134, 3), dtype=np.uint8), (9, 9), 3)
    yy, xx = np.mgrid[0:75, 0:101]
    ramp = np.stack([(xx * 2.5) % 256, (yy * 3.3) % 256, ((xx + yy) * 1.7) % 256], -1).astype(np.uint8)
    noise = rng.integers(0, 256, (37, 51, 3), dtype=np.uint8)
    sf = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
          "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, "440": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440,
          "411": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411}
    files = []
    for k, v in sf.items():
        img = {"444": smooth, "422": ramp, "420": smooth, "440": ramp, "411": noise}[k]
        p = os.path.join(d, f"s{k}.jpg")
        cv2.imwrite(p, img, [cv2.IMWRITE_JPEG_QUALITY, 60 if k != "411" else 92, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, v])
        files.append(p)
    p = os.path.join(d, "rst3.jpg"); cv2.imwrite(p, smooth, [cv2.IMWRITE_JPEG_RST_INTERVAL, 3, cv2.IMWRITE_JPEG_QUALITY, 75]); files.append(p)
    p = os.path.join(d, "grey.jpg"); cv2.imwrite(p, cv2.cvtColor(ramp, cv2.COLOR_BGR2GRAY)); files.append(p)
    p = os.path.join(d, "opt.jpg"); cv2.imwrite(p, noise, [cv2.IMWRITE_JPEG_OPTIMIZE, 1, cv2.IMWRITE_JPEG_QUALITY, 50]); files.append(p)
    for name, img, extra in (("prog420", smooth, [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf["420"], cv2.IMWRITE_JPEG_QUALITY, 70]),
                             ("prog444_rst", ramp, [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf["444"], cv2.IMWRITE_JPEG_RST_INTERVAL, 2]),
                             ("prog_grey", cv2.cvtColor(noise, cv2.COLOR_BGR2GRAY), [cv2.IMWRITE_JPEG_QUALITY, 40])):
        p = os.path.join(d, name + ".jpg")           # progressive: spectral selection + successive approximation scans
        cv2.imwrite(p, img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1] + extra)
        files.append(p)
    for w, h in ((1, 1), (7, 3), (16, 17)):
        p = os.path.join(d, f"tiny_{w}x{h}.jpg"); cv2.imwrite(p, rng.integers(0, 256, (h, w, 3), dtype=np.uint8)); files.append(p)
    # the frame the GPU side-by-side test feeds both libraries: non-square, letterboxed by network_predict_image
    p = os.path.join(d, "frame_500x300.jpg")
    cv2.imwrite(p, rng.integers(0, 256, (300, 500, 3), dtype=np.uint8), [cv2.IMWRITE_JPEG_QUALITY, 95]); files.append(p)
    out = {}
    for p in files:
        im = ref.load_image_color(p.encode(), 0, 0)
        a = np.ctypeslib.as_array(im.data, shape=(im.c, im.h, im.w)).copy()
        ref.free_image(im)
        b = np.rint(a * 255.0).astype(np.uint8)
        assert np.array_equal((b.astype(np.float64) / 255.0).astype(np.float32), a)     # the library's float is byte / 255.
        out[os.path.basename(p)] = b
    np.savez_compressed(os.path.join(GOLD, "jpeg_cases.npz"), **out)
    print(f"[jpeg] {len(files)} fixtures, {sum(os.path.getsize(p) for p in files)} bytes; reference outputs written")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    what = sys.argv[1:] or ["decode", "darknet", "keras", "tracker", "resize", "heatmap", "jpeg"]
    if "decode" in what:
        make_decode_goldens()
    if "darknet" in what:
        make_darknet_golden()
    if "keras" in what:
        make_keras_golden()
    if "tracker" in what:
        make_tracker_golden()
    if "resize" in what:
        make_resize_golden()
    if "heatmap" in what:
        make_heatmap_golden()
    if "jpeg" in what:
        make_jpeg_golden()
