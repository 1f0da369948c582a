"""GPU test of the reference's EXISTING C-ABI (include/darknet_compat.h): the ctypes call sequence of
models_detection/YOLO.py:124-170 against libb200track.so, checked against committed outputs of the reference's
own libdarknet.so and -- where oracle/_ref/libdarknet.so travelled -- against that library run side by side."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

from oracle import darknet_oracle, darknet_ref
from object_tracking_b200 import _native, weights as W

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


class BOX(C.Structure):            # YOLO.py:6-10
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("w", C.c_float), ("h", C.c_float)]


class DETECTION(C.Structure):      # YOLO.py:12-18
    _fields_ = [("bbox", BOX), ("classes", C.c_int), ("prob", C.POINTER(C.c_float)), ("mask", C.POINTER(C.c_float)),
                ("objectness", C.c_float), ("sort_class", C.c_int)]


class IMAGE(C.Structure):          # YOLO.py:20-24
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("c", C.c_int), ("data", C.POINTER(C.c_float))]


class METADATA(C.Structure):       # YOLO.py:26-28
    _fields_ = [("classes", C.c_int), ("names", C.POINTER(C.c_char_p))]


class FEATURE(C.Structure):        # YOLO.py:30-32
    _fields_ = [("size", C.c_int), ("feat", C.POINTER(C.c_float))]


class DIMS(C.Structure):           # YOLO.py:34-37
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("c", C.c_int)]


def bind(lib):
    """The argtypes/restypes exactly as YOLO.py:60-119 sets them."""
    lib.network_width.argtypes = [C.c_void_p]; lib.network_width.restype = C.c_int
    lib.network_height.argtypes = [C.c_void_p]; lib.network_height.restype = C.c_int
    lib.cuda_set_device.argtypes = [C.c_int]
    lib.get_network_boxes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]
    lib.get_network_boxes.restype = C.POINTER(DETECTION)
    lib.free_detections.argtypes = [C.POINTER(DETECTION), C.c_int]
    lib.load_network.argtypes = [C.c_char_p, C.c_char_p, C.c_int]; lib.load_network.restype = C.c_void_p
    lib.do_nms_obj.argtypes = [C.POINTER(DETECTION), C.c_int, C.c_int, C.c_float]
    lib.free_image.argtypes = [IMAGE]
    lib.get_metadata.argtypes = [C.c_char_p]; lib.get_metadata.restype = METADATA
    lib.load_image_color.argtypes = [C.c_char_p, C.c_int, C.c_int]; lib.load_image_color.restype = IMAGE
    lib.rgbgr_image.argtypes = [IMAGE]
    lib.network_predict_image.argtypes = [C.c_void_p, IMAGE]; lib.network_predict_image.restype = C.POINTER(C.c_float)
    lib.network_extract_feat.argtypes = [C.c_void_p, C.c_int]; lib.network_extract_feat.restype = FEATURE
    lib.layer_dims.argtypes = [C.c_void_p, C.c_int]; lib.layer_dims.restype = DIMS
    return lib


def detect(lib, net, meta, im, thresh=.5, hier=.5, nms=.45):
    """YOLO.detect (YOLO.py:140-162) on an IMAGE already in memory."""
    num = C.c_int(0)
    lib.network_predict_image(net, im)
    dets = lib.get_network_boxes(net, im.w, im.h, thresh, hier, None, 0, C.byref(num))
    n = num.value
    if nms:
        lib.do_nms_obj(dets, n, meta.classes, nms)
    res = []
    for j in range(n):
        for i in range(meta.classes):
            if dets[j].prob[i] > 0:
                b = dets[j].bbox
                res.append((meta.names[i].decode(), dets[j].prob[i], (b.x, b.y, b.w, b.h)))
    res = sorted(res, key=lambda x: -x[1])
    lib.free_detections(dets, n)
    return res


@pytest.fixture(scope="module")
def files():
    d = tempfile.mkdtemp()
    w = W.synthetic_yolo_weights(80, seed=0)
    cfg, wts = os.path.join(d, "yolov2.cfg"), os.path.join(d, "synthetic.weights")
    darknet_ref.write_yolov2_cfg(cfg, 80, 416)
    W.write_darknet_weights(wts, w, 80)
    names = os.path.join(d, "coco.names")
    from object_tracking_b200.models_detection._common import COCO_NAMES
    open(names, "w").write("\n".join(COCO_NAMES) + "\n")
    data = os.path.join(d, "coco.data")
    open(data, "w").write(f"classes= 80\nnames = {names}\n")
    return cfg, wts, data


def as_image(frame_u8):
    chw = np.ascontiguousarray(np.transpose(frame_u8.astype(np.float32) / np.float32(255.), (2, 0, 1)))
    im = IMAGE(frame_u8.shape[1], frame_u8.shape[0], 3, chw.ctypes.data_as(C.POINTER(C.c_float)))
    im._keep = chw
    return im


def test_yolo_py_call_sequence_matches_libdarknet_golden(files):
    cfg, wts, data = files
    z = np.load(os.path.join(GOLD, "darknet_416.npz"))
    lib = bind(C.CDLL(_native.LIB_PATH))
    lib.cuda_set_device(0)
    net = lib.load_network(cfg.encode(), wts.encode(), 0)
    assert net
    meta = lib.get_metadata(data.encode())
    assert meta.classes == 80 and meta.names[0] == b"person" and meta.names[79] == b"toothbrush"
    assert lib.network_width(net) == 416 and lib.network_height(net) == 416
    d = lib.layer_dims(net, 25)
    assert (d.h, d.w, d.c) == (13, 13, 1024)                         # YOLO.get_layer_dims -> (h, w, c)
    frame = np.random.default_rng(int(z["frame_seed"])).integers(0, 256, (416, 416, 3), dtype=np.uint8)
    res = detect(lib, net, meta, as_image(frame))
    ref = darknet_oracle.yolo_detect_list(z["det_boxes"], z["det_obj"], z["det_prob"], [meta.names[i].decode() for i in range(80)])
    assert len(res) == len(ref) > 0
    for (n1, p1, b1), (n2, p2, b2) in zip(res, ref):
        assert n1 == n2 and abs(p1 - p2) < 1e-4
        assert np.abs(np.array(b1) - np.array(b2)).max() < 2e-2      # pixels
    f = lib.network_extract_feat(net, 25)                            # YOLO.extract: flat CHW
    assert f.size == 13 * 13 * 1024
    feat = np.ctypeslib.as_array(f.feat, shape=(f.size,)).reshape(1024, 13, 13)
    assert np.abs(feat.reshape(1024, -1).max(1) - z["feat_globalmax"]).max() < 3e-3
    assert np.abs(feat[::16] - z["feat_sub"]).max() < 3e-3
    # no NMS requested (config nms = 0): thresholded detections only
    num = C.c_int(0)
    dets = lib.get_network_boxes(net, 416, 416, .5, .5, None, 0, C.byref(num))
    assert num.value == 13 * 13 * 5
    n_pos = sum(1 for j in range(num.value) if dets[j].objectness > 0)
    assert n_pos >= len(ref)
    lib.free_detections(dets, num.value)


def test_errors_do_not_exit_the_process(files):
    lib = bind(C.CDLL(_native.LIB_PATH))
    assert not lib.load_network(b"/nonexistent.cfg", b"x.weights", 0)
    assert b"cannot read cfg" in _native.lib().b2t_last_error()
    cfg, wts, data = files
    assert not lib.load_network(cfg.encode(), b"/nonexistent.weights", 0)
    im = lib.load_image_color(b"/nonexistent.jpg", 0, 0)
    assert not im.data


@pytest.mark.skipif(not darknet_ref.available(), reason="oracle/_ref/libdarknet.so not present")
def test_letterbox_and_detections_side_by_side_with_reference_library(files):
    """Same cfg/weights/image through the reference's libdarknet.so (CPU) and libb200track.so (B200)."""
    cfg, wts, data = files
    ref = bind(C.CDLL(darknet_ref.LIB_PATH))
    ours = bind(C.CDLL(_native.LIB_PATH))
    ref.load_network.restype = C.c_void_p
    cwd = os.getcwd()
    net_r = ref.load_network(cfg.encode(), wts.encode(), 0)
    net_o = ours.load_network(cfg.encode(), wts.encode(), 0)
    os.chdir(cwd)
    meta = ours.get_metadata(data.encode())
    rng = np.random.default_rng(77)
    for (h, w) in ((416, 416), (300, 500), (480, 270)):
        # smooth-ish image so that bilinear letterboxing matters but thresholds are not razor-edge
        frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        a = detect(ref, net_r, meta, as_image(frame))
        b = detect(ours, net_o, meta, as_image(frame))
        fr = ref.network_extract_feat(net_r, 25)
        fo = ours.network_extract_feat(net_o, 25)
        fa = np.ctypeslib.as_array(fr.feat, shape=(fr.size,)).copy()
        fb = np.ctypeslib.as_array(fo.feat, shape=(fo.size,)).copy()
        assert np.abs(fa - fb).max() < 5e-3 * max(1.0, np.abs(fa).max() / 10), (h, w)
        assert [x[0] for x in a] == [x[0] for x in b], (h, w)
        for (n1, p1, b1), (n2, p2, b2) in zip(a, b):
            assert abs(p1 - p2) < 2e-4 and np.abs(np.array(b1) - np.array(b2)).max() < 5e-2


@pytest.mark.skipif(not darknet_ref.available(), reason="oracle/_ref/libdarknet.so not present")
def test_yolo_detect_on_a_jpeg_path_side_by_side(files):
    """the reference's own call sequence on a .jpg PATH (YOLO.py:140-162: load_image_color ->
    network_predict_image -> get_network_boxes -> do_nms_obj -> free_image), non-square frame, through both
    libraries; then the Python plugin YOLO.detect(path) against the same reference result.  Planted detector weights
    and thresh = 0.25 so that the letterboxed frame carries a dozen detections."""
    cfg, _, data = files
    wd = W.synthetic_detector_weights(80, seed=0)
    wts = os.path.join(os.path.dirname(cfg), "planted.weights")
    W.write_darknet_weights(wts, wd, 80)
    ref = bind(C.CDLL(darknet_ref.LIB_PATH))
    ours = bind(C.CDLL(_native.LIB_PATH))
    cwd = os.getcwd()
    net_r = ref.load_network(cfg.encode(), wts.encode(), 0)
    net_o = ours.load_network(cfg.encode(), wts.encode(), 0)
    os.chdir(cwd)
    meta = ours.get_metadata(data.encode())
    path = os.path.join(GOLD, "jpeg", "frame_500x300.jpg")
    out = []
    for lib, net in ((ref, net_r), (ours, net_o)):
        im = lib.load_image_color(path.encode(), 0, 0)
        assert im.data and (im.w, im.h, im.c) == (500, 300, 3)
        out.append(detect(lib, net, meta, im, thresh=.25))
        lib.free_image(im)
    a, b = out
    assert len(a) >= 5 and [x[0] for x in a] == [x[0] for x in b]
    for (n1, p1, b1), (n2, p2, b2) in zip(a, b):
        assert abs(p1 - p2) < 2e-4 and np.abs(np.array(b1) - np.array(b2)).max() < 5e-2      # pixels of a 500x300 frame
    # the plugin class on the same path: cv2 decodes the file (libjpeg's chroma upsampling is not bit-identical to the
    # reference decoder's), BGR -> RGB, letterbox on the device, boxes un-mapped to the 500x300 frame
    import copy
    from object_tracking_b200.models_detection.YOLO import YOLO
    from object_tracking_b200.models_detection._common import load_config
    conf = copy.deepcopy(load_config(None))
    conf["model_detector"]["thresh"] = 0.25
    y = YOLO(config=conf, weights=wd, max_batch=1)
    c = y.detect(path)
    assert len(c) >= 5
    hits = 0
    for n, p, bx in c:
        for n2, p2, bx2 in a:
            if n == n2 and abs(p - p2) < 0.03 and np.abs(np.array(bx) - np.array(bx2)).max() < 3.0:    # pixels
                hits += 1
                break
    assert hits >= len(c) - 2, (hits, len(c), len(a))


@pytest.mark.skipif(not darknet_ref.available(), reason="oracle/_ref/libdarknet.so not present")
@pytest.mark.parametrize("variant", ["voc", "coco"])
def test_tiny_yolov2_graph_side_by_side(tmp_path, variant):
    """cfg/yolov2-tiny-voc.cfg (20 classes, conv_8 = 1024) and cfg/yolov2-tiny.cfg (80 classes, conv_8 = 512): nine conv
    layers, conv_1 with 16 channels, the sixth maxpool with stride 1 -- through the reference's C library and through
    libb200track.so (compat ABI: one frame = the chain schedule; engine API: batch 3 = one kernel per layer)."""
    import torch
    from object_tracking_b200.engine import DetectorEngine
    n_class, f8, anchors = (20, 1024, darknet_ref.TINY_VOC_ANCHORS) if variant == "voc" else (80, 512, darknet_ref.COCO_ANCHORS)
    d = str(tmp_path)
    cfg, wts = os.path.join(d, "tiny.cfg"), os.path.join(d, "tiny.weights")
    darknet_ref.write_tiny_cfg(cfg, n_class, 416, f8, anchors)
    darknet_ref.write_tiny_weights(wts, n_class, f8, seed=3)
    names = os.path.join(d, "n.names")
    open(names, "w").write("\n".join(f"c{i}" for i in range(n_class)) + "\n")
    data = os.path.join(d, "n.data")
    open(data, "w").write(f"classes= {n_class}\nnames = {names}\n")
    ref = bind(C.CDLL(darknet_ref.LIB_PATH))
    ours = bind(C.CDLL(_native.LIB_PATH))
    cwd = os.getcwd()
    net_r = ref.load_network(cfg.encode(), wts.encode(), 0)
    net_o = ours.load_network(cfg.encode(), wts.encode(), 0)
    os.chdir(cwd)
    assert net_o, _native.lib().b2t_last_error()
    meta = ours.get_metadata(data.encode())
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (3, 416, 416, 3), dtype=np.uint8)
    for k, n_layer in enumerate((13, 14, 15)):                        # darknet layers 12, 13 (features) and 14 (head)
        do, dr = ours.layer_dims(net_o, n_layer), ref.layer_dims(net_r, n_layer)
        assert (do.w, do.h, do.c) == (dr.w, dr.h, dr.c), n_layer
    thr = .2 if variant == "voc" else .05
    a = detect(ref, net_r, meta, as_image(frames[0]), thresh=thr)
    b = detect(ours, net_o, meta, as_image(frames[0]), thresh=thr)
    for n_layer, tol in ((12, 2e-3), (14, 3e-3), (15, 3e-3)):         # pool_6 (the stride-1 max-pool), conv_8, head logits
        fr, fo = ref.network_extract_feat(net_r, n_layer), ours.network_extract_feat(net_o, n_layer)
        assert fr.size == fo.size, n_layer
        fa = np.ctypeslib.as_array(fr.feat, shape=(fr.size,)).copy()
        fb = np.ctypeslib.as_array(fo.feat, shape=(fo.size,)).copy()
        assert np.abs(fa - fb).max() < tol * max(1.0, np.abs(fa).max() / 10), (n_layer, np.abs(fa - fb).max(), np.abs(fa).max())
    logits_ref = fa.reshape(5 * (5 + n_class), 13, 13)
    assert (len(a) >= 3 or variant == "coco") and [x[0] for x in a] == [x[0] for x in b]
    for (n1, p1, b1), (n2, p2, b2) in zip(a, b):
        assert abs(p1 - p2) < 3e-4 and np.abs(np.array(b1) - np.array(b2)).max() < 5e-2
    # the engine API at batch 3 (one kernel per layer) gives the same logits for frame 0 as the library
    e = DetectorEngine(n_class=n_class, max_batch=3, semantics="darknet", graph="tiny", tiny_filters=f8)
    e.load_darknet_weights(wts)
    e.finalize()
    lg = e.forward(torch.from_numpy(frames).cuda()).cpu().numpy()
    got = lg[0].reshape(13, 13, -1).transpose(2, 0, 1)
    assert np.abs(got - logits_ref).max() < 3e-3 * max(1.0, np.abs(logits_ref).max() / 10)
    one = e.forward(torch.from_numpy(frames[:1]).cuda()).cpu().numpy()   # batch 1: chain schedule
    assert np.abs(one[0] - lg[0]).max() < 5e-4
