"""ctypes driver for the REFERENCE darknet C library built CPU-only into oracle/_ref/libdarknet.so
(recipe: oracle/Makefile).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Call sequence = models_detection/YOLO.py:124-170 of the reference (load_network ->
network_predict -> get_network_boxes -> do_nms_obj -> network_extract_feat / layer_dims), with
``network_predict`` on a float CHW buffer instead of ``network_predict_image`` so no image file
is needed.  Struct layouts mirror darknet/include/darknet.h:507-525 and darknet/src/network.h:11-20.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libdarknet.so")


class BOX(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("w", C.c_float), ("h", C.c_float)]


class DETECTION(C.Structure):
    _fields_ = [("bbox", BOX), ("classes", C.c_int), ("prob", C.POINTER(C.c_float)),
                ("mask", C.POINTER(C.c_float)), ("objectness", C.c_float), ("sort_class", C.c_int)]


class FEATURE(C.Structure):
    _fields_ = [("size", C.c_int), ("feat", C.POINTER(C.c_float))]


class DIMS(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("c", C.c_int)]


def available() -> bool:
    return os.path.exists(LIB_PATH)


def write_yolov2_cfg(path: str, n_class: int = 80, size: int = 416) -> None:
    """Emit a darknet cfg describing the same graph as cfg/yolov2.cfg (31 layers, region head)
    generated from our own layer table, so the library can be driven where /root/reference is
    absent (GPU box)."""
    t = [(32, 3, 1), (64, 3, 1), (128, 3, 0), (64, 1, 0), (128, 3, 1), (256, 3, 0), (128, 1, 0), (256, 3, 1),
         (512, 3, 0), (256, 1, 0), (512, 3, 0), (256, 1, 0), (512, 3, 1), (1024, 3, 0), (512, 1, 0),
         (1024, 3, 0), (512, 1, 0), (1024, 3, 0), (1024, 3, 0), (1024, 3, 0)]
    s = [f"[net]\nbatch=1\nsubdivisions=1\nwidth={size}\nheight={size}\nchannels=3\n"]
    conv = "[convolutional]\nbatch_normalize=1\nfilters={f}\nsize={k}\nstride=1\npad=1\nactivation=leaky\n"
    for f, k, pool in t:
        s.append(conv.format(f=f, k=k))
        if pool:
            s.append("[maxpool]\nsize=2\nstride=2\n")
    s.append("[route]\nlayers=-9\n")
    s.append(conv.format(f=64, k=1))
    s.append("[reorg]\nstride=2\n")
    s.append("[route]\nlayers=-1,-4\n")
    s.append(conv.format(f=1024, k=3))
    s.append(f"[convolutional]\nsize=1\nstride=1\npad=1\nfilters={5 * (5 + n_class)}\nactivation=linear\n")
    s.append("[region]\nanchors =  0.57273, 0.677385, 1.87446, 2.06253, 3.33843, 5.47434, 7.88282, 3.52778, "
             f"9.77052, 9.16828\nbias_match=1\nclasses={n_class}\ncoords=4\nnum=5\nsoftmax=1\njitter=.3\n"
             "rescore=1\nobject_scale=5\nnoobject_scale=1\nclass_scale=1\ncoord_scale=1\nabsolute=1\n"
             "thresh = .6\nrandom=1\n")
    with open(path, "w") as f:
        f.write("\n".join(s))


class DarknetRef:
    def __init__(self, cfg_path: str, weights_path: str):
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle` where /root/reference exists")
        lib = C.CDLL(LIB_PATH, C.RTLD_GLOBAL)
        lib.load_network.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        lib.load_network.restype = C.c_void_p
        lib.network_predict.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        lib.network_predict.restype = C.POINTER(C.c_float)
        lib.network_width.argtypes = [C.c_void_p]
        lib.network_width.restype = C.c_int
        lib.network_height.argtypes = [C.c_void_p]
        lib.network_height.restype = C.c_int
        lib.get_network_boxes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float,
                                          C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]
        lib.get_network_boxes.restype = C.POINTER(DETECTION)
        lib.do_nms_obj.argtypes = [C.POINTER(DETECTION), C.c_int, C.c_int, C.c_float]
        lib.free_detections.argtypes = [C.POINTER(DETECTION), C.c_int]
        lib.network_extract_feat.argtypes = [C.c_void_p, C.c_int]
        lib.network_extract_feat.restype = FEATURE
        lib.layer_dims.argtypes = [C.c_void_p, C.c_int]
        lib.layer_dims.restype = DIMS
        self.lib = lib
        self.net = lib.load_network(cfg_path.encode(), weights_path.encode(), 0)
        self.w, self.h = lib.network_width(self.net), lib.network_height(self.net)

    def predict(self, chw: np.ndarray) -> None:
        x = np.ascontiguousarray(chw, dtype=np.float32)
        assert x.size == 3 * self.w * self.h
        self._keep = x
        self.lib.network_predict(self.net, x.ctypes.data_as(C.POINTER(C.c_float)))

    def layer_dims(self, n: int) -> Tuple[int, int, int]:
        d = self.lib.layer_dims(self.net, n)
        return d.h, d.w, d.c                       # YOLO.py:136-138 order

    def extract(self, n: int) -> np.ndarray:
        f = self.lib.network_extract_feat(self.net, n)
        return np.ctypeslib.as_array(f.feat, shape=(f.size,)).copy()

    def detect(self, im_w: int, im_h: int, thresh: float, hier: float, nms: float, n_class: int,
               ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """-> (boxes (n,4) cx,cy,w,h pixels, objectness (n,), prob (n,classes)) after do_nms_obj,
        in the library's post-NMS array order (YOLO.py:144-158 reads exactly this)."""
        num = C.c_int(0)
        dets = self.lib.get_network_boxes(self.net, im_w, im_h, thresh, hier, None, 0, C.byref(num))
        n = num.value
        if nms:
            self.lib.do_nms_obj(dets, n, n_class, nms)
        boxes = np.zeros((n, 4), np.float32)
        obj = np.zeros(n, np.float32)
        prob = np.zeros((n, n_class), np.float32)
        for j in range(n):
            d = dets[j]
            boxes[j] = (d.bbox.x, d.bbox.y, d.bbox.w, d.bbox.h)
            obj[j] = d.objectness
            prob[j] = np.ctypeslib.as_array(d.prob, shape=(n_class,))
        self.lib.free_detections(dets, n)
        return boxes, obj, prob


TINY_VOC_ANCHORS = "1.08,1.19,  3.42,4.41,  6.63,11.38,  9.42,5.11,  16.62,10.52"
COCO_ANCHORS = "0.57273, 0.677385, 1.87446, 2.06253, 3.33843, 5.47434, 7.88282, 3.52778, 9.77052, 9.16828"


def write_tiny_cfg(path: str, n_class: int, size: int = 416, f8: int = 1024, anchors: str = TINY_VOC_ANCHORS) -> None:
    """The layer list of darknet/cfg/yolov2-tiny-voc.cfg (f8 = 1024, 20 classes) / yolov2-tiny.cfg (f8 = 512, 80
    classes): six conv + maxpool pairs (the sixth pool has stride 1), two more 3x3 convs, the 1x1 head, region."""
    conv = "[convolutional]\nbatch_normalize=1\nfilters={f}\nsize=3\nstride=1\npad=1\nactivation=leaky\n"
    s = [f"[net]\nbatch=1\nsubdivisions=1\nwidth={size}\nheight={size}\nchannels=3\nmomentum=0.9\ndecay=0.0005\n"]
    for i, f in enumerate((16, 32, 64, 128, 256, 512)):
        s.append(conv.format(f=f))
        s.append("[maxpool]\nsize=2\nstride=%d\n" % (1 if i == 5 else 2))
    s.append(conv.format(f=1024))
    s.append(conv.format(f=f8))
    s.append(f"[convolutional]\nsize=1\nstride=1\npad=1\nfilters={5 * (5 + n_class)}\nactivation=linear\n")
    s.append(f"[region]\nanchors = {anchors}\nbias_match=1\nclasses={n_class}\ncoords=4\nnum=5\nsoftmax=1\njitter=.2\n"
             "rescore=1\nobject_scale=5\nnoobject_scale=1\nclass_scale=1\ncoord_scale=1\nabsolute=1\nthresh = .6\nrandom=1\n")
    with open(path, "w") as f:
        f.write("\n".join(s))


def write_tiny_weights(path: str, n_class: int, f8: int = 1024, seed: int = 0) -> None:
    """Seeded random-init weights for the tiny graph in darknet's file order (parser.c:1149-1198): per conv
    biases(beta) [scales(gamma) rolling_mean rolling_variance] weights[Cout][Cin][kh][kw]; v0.1 header."""
    import struct
    rng = np.random.default_rng(seed)
    chans = [(3, 16, 3), (16, 32, 3), (32, 64, 3), (64, 128, 3), (128, 256, 3), (256, 512, 3), (512, 1024, 3), (1024, f8, 3),
             (f8, 5 * (5 + n_class), 1)]
    with open(path, "wb") as f:
        f.write(struct.pack("<iiii", 0, 1, 0, 0))
        for i, (ci, co, k) in enumerate(chans):
            last = i == len(chans) - 1
            if last:
                b = (0.1 * rng.standard_normal(co)).astype("<f4")
                b.reshape(5, 5 + n_class)[:, 4] -= 1.0
                f.write(b.tobytes())
            else:
                f.write((0.1 * rng.standard_normal(co)).astype("<f4").tobytes())          # beta
                f.write(rng.uniform(0.5, 1.5, co).astype("<f4").tobytes())                # gamma
                f.write((0.1 * rng.standard_normal(co)).astype("<f4").tobytes())          # mean
                f.write(rng.uniform(0.5, 1.5, co).astype("<f4").tobytes())                # var
            gain = (0.3 if last else 1.0) * np.sqrt(2.0 / (k * k * ci))
            f.write((rng.standard_normal((co, ci, k, k)) * gain).astype("<f4").tobytes())
