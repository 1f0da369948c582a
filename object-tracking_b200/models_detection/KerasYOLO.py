"""Drop-in for the inference half of the reference's ``models_detection/KerasYOLO.py``.

Kept surface: class constants (:37-59), ``KerasYOLO(argv)`` with the 6-key override dict (:67-79),
``predict(input_path, output_path)`` (:522-537), ``extract(input_path, layer)`` (:509-520),
``normalize_input`` (:412-413), ``load_weights``; plus array-in / boxes-out batch variants for callers
that already hold frames.  ``model.predict`` (the TensorFlow session call, :531) is replaced by
``DetectorEngine.forward`` and ``decode_netout`` by the device kernel.  Training (loss_fxn/train) is out of
scope.
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np
import torch

from ..engine import DetectorEngine, rows_to_host
from ..utility.utils import BoundBox, boxes_from_rows, draw_boxes, normalize
from ..weights import ANCHORS, read_darknet_weights, synthetic_yolo_weights
from ._common import load_frame

LABELS_COCO = [
    'person', 'bicycle', 'car', 'motorcycle', 'airplane', 'bus', 'train', 'truck', 'boat', 'traffic light',
    'fire hydrant', 'stop sign', 'parking meter', 'bench', 'bird', 'cat', 'dog', 'horse', 'sheep', 'cow',
    'elephant', 'bear', 'zebra', 'giraffe', 'backpack', 'umbrella', 'handbag', 'tie', 'suitcase', 'frisbee',
    'skis', 'snowboard', 'sports ball', 'kite', 'baseball bat', 'baseball glove', 'skateboard', 'surfboard',
    'tennis racket', 'bottle', 'wine glass', 'cup', 'fork', 'knife', 'spoon', 'bowl', 'banana', 'apple',
    'sandwich', 'orange', 'broccoli', 'carrot', 'hot dog', 'pizza', 'donut', 'cake', 'chair', 'couch',
    'potted plant', 'bed', 'dining table', 'toilet', 'tv', 'laptop', 'mouse', 'remote', 'keyboard',
    'cell phone', 'microwave', 'oven', 'toaster', 'sink', 'refrigerator', 'book', 'clock', 'vase', 'scissors',
    'teddy bear', 'hair drier', 'toothbrush']


class KerasYOLO:
    LABELS = LABELS_COCO
    IMAGE_H, IMAGE_W = 416, 416
    GRID_H, GRID_W = 13, 13
    BOX = 5
    CLASS = len(LABELS)
    CLASS_WEIGHTS = np.ones(CLASS, dtype='float32')
    OBJ_THRESHOLD = 0.5
    NMS_THRESHOLD = 0.45
    ANCHORS = list(ANCHORS)
    BATCH_SIZE = 32
    TRUE_BOX_BUFFER = 50
    MAX_BOX_PER_IMAGE = 50
    weight_path = 'darknet/yolov2.weights'
    model = None

    def __init__(self, argv={}, device: int = 0, convlstm_units: int = 0, weights: Optional[dict] = None,
                 keep_prepool: bool = False):
        if len(argv) == 6:
            self.LABELS = argv['LABELS']
            self.CLASS = len(self.LABELS)
            self.CLASS_WEIGHTS = np.ones(self.CLASS, dtype='float32')
            self.BATCH_SIZE = argv['BATCH_SIZE']
            self.IMAGE_H, self.IMAGE_W = argv['IMAGE_H'], argv['IMAGE_W']
            self.GRID_H, self.GRID_W = argv['GRID_H'], argv['GRID_W']
        if self.IMAGE_H != 32 * self.GRID_H or self.IMAGE_W != 32 * self.GRID_W:
            raise ValueError("IMAGE_H/W must be 32 * GRID_H/W (five 2x2 max-pools)")
        self._device, self._convlstm_units, self._weights, self._keep = device, convlstm_units, weights, keep_prepool
        self.load_model()

    def load_model(self):
        """KerasYOLO.py:239-407 (graph) + init_weights (:244-274)."""
        self.model = DetectorEngine(n_class=self.CLASS, image_size=self.IMAGE_H, max_batch=max(1, self.BATCH_SIZE),
                                    semantics="keras", bn_eps=1e-3, device=self._device,
                                    convlstm_units=self._convlstm_units, keep_prepool=self._keep)
        if self._weights is not None:
            self.model.set_weights(self._weights)
            self.synthetic_weights = False
        elif os.path.exists(self.weight_path):
            self.model.set_weights(read_darknet_weights(self.weight_path, self.CLASS))
            self.synthetic_weights = False
        else:
            self.model.set_weights(synthetic_yolo_weights(self.CLASS, seed=0))
            self.synthetic_weights = True
        if not self._convlstm_units:
            self.model.finalize()

    def load_weights(self, path):
        self.model.set_weights(read_darknet_weights(path, self.CLASS))
        self.model.finalize()

    def normalize_input(self, image):
        return normalize(image)

    # ------------------------------------------------------------------ batch API (arrays in, boxes out)
    def predict_batch(self, frames) -> List[List[BoundBox]]:
        """frames: (B,H,W,3) uint8 numpy/tensor (colour order as the caller has it, KerasYOLO.py:525-528)."""
        t = torch.as_tensor(np.ascontiguousarray(frames) if isinstance(frames, np.ndarray) else frames)
        t = t.to(self.model.device).contiguous()
        logits = self.model.forward(t)
        boxes, counts = self.model.decode(logits, self.OBJ_THRESHOLD, self.NMS_THRESHOLD, self.ANCHORS)
        return [boxes_from_rows(r, self.CLASS) for r in rows_to_host(boxes, counts)]

    # ------------------------------------------------------------------ reference API
    def _load(self, input_path):
        """cv2.imread on the host (KerasYOLO.py:525), cv2.resize on the device (:526, engine.resize_frames --
        bit-identical to OpenCV's INTER_LINEAR); returns (original HWC uint8 array, resized CUDA tensor)."""
        image = load_frame(input_path)
        t = torch.from_numpy(np.ascontiguousarray(image[None])).to(self.model.device)
        if image.shape[0] == self.IMAGE_H and image.shape[1] == self.IMAGE_W:
            return image, t[0]
        return image, self.model.resize_frames(t, self.IMAGE_H)[0]

    def extract(self, input_path, layer):
        _, resized = self._load(input_path)
        self.model.forward(resized[None].contiguous())
        return self.model.extract(layer, 1)[0].cpu().numpy()

    def predict(self, input_path, output_path):
        image, resized = self._load(input_path)
        boxes = self.predict_batch(resized[None])[0]
        image = draw_boxes(image, boxes, self.LABELS)
        print(len(boxes), 'Bounding Boxes Found')
        print("File Saved to", output_path)
        import cv2
        cv2.imwrite(output_path, image)
        return boxes
