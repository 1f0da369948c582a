"""Drop-in for the reference's ``models_detection/YOLO.py`` (the ctypes wrapper around libdarknet.so),
with the darknet forward / region layer / NMS running on the B200 through libb200track.so.

Kept surface (YOLO.py:38-180): ``YOLO([cpu_mode, gpu_id])``, attributes ``CLASSES, NMS, THRESH, HIER_THRESH,
META, CONFIG, WEIGHTS, gpu_id, cpu_mode``, ``get_layer_dims(n) -> (h,w,c)``, ``detect(image) ->
[(name, prob, (cx,cy,w,h))] sorted by -prob``, ``extract(n) -> flat CHW float64``,
``extract_spatio_info(frame, layer) -> (detections of config classes, feature)``.
Darknet semantics (BN 1/(sqrt(var)+1e-6), reorg ordering, per-anchor softmax, objectness NMS) are selected
with ``semantics="darknet"`` in the engine.  ``image`` may be a path or an HWC uint8 RGB array; frames that
are not net-sized go through darknet's letterbox ingest on the device (engine.letterbox_frames).
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import numpy as np
import torch

from ..engine import DetectorEngine, rows_to_host
from ..weights import ANCHORS, synthetic_detector_weights
from ._common import COCO_NAMES, load_config, load_frame

# darknet cfg/yolov2.cfg layer index -> tensor name kept by the engine
_LAYER_NAMES = {0: "norm_1", 1: "pool_1", 2: "norm_2", 3: "pool_2", 4: "norm_3", 5: "norm_4", 6: "norm_5",
                7: "pool_5", 8: "norm_6", 9: "norm_7", 10: "norm_8", 11: "pool_8", 12: "norm_9", 13: "norm_10",
                14: "norm_11", 15: "norm_12", 16: "norm_13", 17: "pool_13", 18: "norm_14", 19: "norm_15",
                20: "norm_16", 21: "norm_17", 22: "norm_18", 23: "norm_19", 24: "norm_20", 25: "norm_13",
                28: "concat", 29: "norm_22", 30: "conv_23"}


class YOLO:
    def __init__(self, argvs=[], config=None, n_class: int = 80, image_size: int = 416, max_batch: int = 4,
                 weights: Optional[dict] = None, names: Optional[List[str]] = None, broadcast: bool = False,
                 rank: int = 0):
        self.config = load_config(config)
        self.gpu_id = self.config["train"]["dgpu_id"]
        self.cpu_mode = self.config["train"]["cpu_only"]
        self.CLASSES = [s.lower() for s in self.config["train"]["classes"]]
        self.NMS = self.config["model_detector"]["nms"]
        self.THRESH = self.config["model_detector"]["thresh"]
        self.HIER_THRESH = self.config["model_detector"]["hier_thresh"]
        self.META = self.config["model_detector"]["meta_file"]
        self.CONFIG = self.config["model_detector"]["config_file"]
        self.WEIGHTS = self.config["model_detector"]["weights_file"]
        self.n_class, self.image_size, self.max_batch = n_class, image_size, max_batch
        self.names = list(names or COCO_NAMES[:n_class])
        self._weights = weights
        self._broadcast, self._rank = broadcast, rank
        self.argv_parser(argvs)
        self.load_detection_model()

    def argv_parser(self, argvs):
        if len(argvs) >= 2:
            self.cpu_mode, self.gpu_id = argvs[0], argvs[1]

    def load_detection_model(self):
        """YOLO.py:128-134 (load_network + get_metadata).  cpu_mode is ignored: this build has no CPU path.
        ``config.model_detector.config_file`` selects the graph by name like the cfg files it points at:
        ``yolov2-tiny-voc.cfg`` / ``yolov2-tiny.cfg`` -> the 9-conv tiny graph (weights file required), anything else
        -> the 23-conv YOLOv2 graph."""
        dev = self.gpu_id if self.gpu_id < torch.cuda.device_count() else 0
        cfg_name = os.path.basename(str(self.CONFIG)).lower()
        tiny = "tiny" in cfg_name
        self.engine = DetectorEngine(n_class=self.n_class, image_size=self.image_size, max_batch=self.max_batch,
                                     semantics="darknet", device=dev, graph="tiny" if tiny else "yolov2",
                                     tiny_filters=1024 if "voc" in cfg_name else 512)
        wpath = os.path.join("darknet", self.WEIGHTS)
        if tiny and self._weights is None and not os.path.exists(wpath) and not (self._broadcast and self._rank != 0):
            raise FileNotFoundError(f"{wpath}: the tiny graph has no synthetic weights; provide the darknet weights file")
        if self._broadcast and self._rank != 0:
            # multi-GPU: rank 0 packs and uploads, everyone else receives the packed blob over NCCL/NVLink
            self.synthetic_weights = self._weights is None and not os.path.exists(wpath)
            self.engine.finalize(upload=False)
        elif self._weights is not None:
            self.engine.set_weights(self._weights)
            self.synthetic_weights = False
        elif os.path.exists(wpath):
            self.engine.load_darknet_weights(wpath)
            self.synthetic_weights = False
        else:                                   # the reference ships no weights: random init
            self.engine.set_weights(synthetic_detector_weights(self.n_class, seed=0))
            self.synthetic_weights = True
        if not (self._broadcast and self._rank != 0):
            self.engine.finalize()
        if self._broadcast:
            self.engine.broadcast_weights(src=0)
        mask = torch.zeros(self.n_class, dtype=torch.uint8)
        for i, n in enumerate(self.names):
            if n in self.CLASSES:
                mask[i] = 1
        self.class_mask = mask.to(self.engine.device)
        self._last_batch = 0

    def get_layer_dims(self, n):
        """network.c:600-607 layer_dims uses layers[n-1]; returns (h, w, c) like YOLO.py:136-138."""
        return self.engine.layer_dims(self._name(n))

    @staticmethod
    def _name(n: int) -> str:
        if (n - 1) not in _LAYER_NAMES:
            raise ValueError(f"darknet layer {n} is not kept by the B200 engine")
        return _LAYER_NAMES[n - 1]

    # ------------------------------------------------------------------ device-side batch API
    def detect_batch(self, frames: torch.Tensor, orig_w: Optional[int] = None, orig_h: Optional[int] = None):
        """frames (B,S,S,3) uint8 RGB on the GPU -> (dets (B,max,8), counts (B)) device tensors, rows
        [cx,cy,w,h px, objectness, prob, class, index] sorted by -prob (YOLO.py:140-162 without the host loop)."""
        logits = self.engine.forward(frames)
        self._last_batch = frames.shape[0]
        return self.engine.region_detect(logits, self.THRESH, self.NMS, orig_w or self.image_size,
                                         orig_h or self.image_size)

    # ------------------------------------------------------------------ reference API
    def detect(self, image):
        """YOLO.py:140-162.  ``image``: a path (decoded on the host; load_image_color gives RGB, so cv2's BGR is
        swapped) or an HWC uint8 RGB array.  Ingest follows the reference: network_predict_image letterboxes the
        frame to the net size (image.c:960-979, device kernel) and get_network_boxes(w, h) un-maps the boxes to
        pixels of the original frame (correct_region_boxes, region_layer.c:336-362)."""
        is_path = isinstance(image, str)
        frame = load_frame(image)
        h, w = frame.shape[:2]
        t = torch.from_numpy(np.ascontiguousarray(frame[None])).to(self.engine.device)
        if (h, w) == (self.image_size, self.image_size):
            if is_path:
                t = t.flip(-1).contiguous()                      # BGR -> RGB; letterbox of a net-sized frame is the identity
        else:
            t = self.engine.letterbox_frames(t, bgr=is_path)
        dets, counts = self.detect_batch(t, w, h)
        rows = rows_to_host(dets, counts)[0]
        return [(self.names[int(r[6])], float(r[5]), (float(r[0]), float(r[1]), float(r[2]), float(r[3]))) for r in rows]

    def extract(self, n):
        """YOLO.py:164-170: the layer's output as a flat CHW float64 vector (frame 0 of the last detect)."""
        x = self.engine.extract(self._name(n), 1)[0]
        return x.permute(2, 0, 1).reshape(-1).cpu().numpy().astype(np.float64)

    def extract_spatio_info(self, frame_path, layer=24):
        out = self.detect(frame_path)
        vis_feat = self.extract(layer)
        return [d for d in out if d[0] in self.CLASSES], vis_feat
