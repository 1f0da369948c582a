"""Drop-in for the inference half of ``models_tracking/MultiObjDetTracker.py``: TimeDistributed YOLOv2 ->
concat[conv_23 logits, conv_feat] -> ConvLSTM2D(512, 3x3, same, return_sequences) -> Conv2D(5*(5+C), 1x1)
-> Reshape -> decode_netout per frame (MultiObjDetTracker.py:160-189, :295-315).  Losses / fit loop are out
of scope.  The ConvLSTM and the head are the same tcgen05 implicit-GEMM kernel as the backbone."""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from ..engine import rows_to_host
from ..models_detection.KerasYOLO import KerasYOLO
from ..models_detection._common import load_frame
from ..utility.utils import BoundBox, boxes_from_rows, draw_boxes
from ..weights import ANCHORS, synthetic_convlstm_weights


class MultiObjDetTracker:
    LABELS_MOT17 = ['1', '2', '3', '4', '5', '6', '7', '8', '9', '10', '11', '12']
    LABELS = LABELS_MOT17
    IMAGE_H, IMAGE_W = 416, 416
    GRID_H, GRID_W = 13, 13
    BOX = 5
    CLASS = len(LABELS)
    OBJ_THRESHOLD = 0.5
    NMS_THRESHOLD = 0.45
    ANCHORS = list(ANCHORS)
    BATCH_SIZE = 1
    SEQUENCE_LENGTH = 4
    MAX_BOX_PER_IMAGE = 50
    CONVLSTM_UNITS = 512
    LOAD_MODEL = False
    model = None
    detector = None
    model_detector = None

    def __init__(self, argv={}, device: int = 0, detector_weights: Optional[dict] = None,
                 tracker_weights: Optional[dict] = None, convlstm_units: Optional[int] = None):
        if 'LABELS' in argv:
            self.LABELS = argv['LABELS']
        self.CLASS = len(self.LABELS)
        if convlstm_units:
            self.CONVLSTM_UNITS = convlstm_units
        argv = dict(argv)
        argv['LABELS'] = self.LABELS
        argv['BATCH_SIZE'] = self.BATCH_SIZE * self.SEQUENCE_LENGTH
        argv['IMAGE_H'], argv['IMAGE_W'] = self.IMAGE_H, self.IMAGE_W
        argv['GRID_H'], argv['GRID_W'] = self.GRID_H, self.GRID_W
        self.detector = KerasYOLO(argv, device=device, convlstm_units=self.CONVLSTM_UNITS, weights=detector_weights)
        self._tracker_weights = tracker_weights
        self.load_model()

    def load_model(self):
        eng = self.detector.model
        n_out = self.BOX * (5 + self.CLASS)
        w = self._tracker_weights or synthetic_convlstm_weights(n_out + 1024, self.CONVLSTM_UNITS, n_out, seed=2)
        eng.set_convlstm_weights(w)
        eng.finalize()
        self.model = self.model_detector = eng
        eng.convlstm_reset()

    def reset(self):
        self.model.convlstm_reset()

    def track_window(self, frames, reset: bool = True):
        """frames (T,H,W,3) uint8 (T <= SEQUENCE_LENGTH): -> (tracker boxes per frame, detector boxes per frame)."""
        eng = self.model
        t = torch.as_tensor(np.ascontiguousarray(frames) if isinstance(frames, np.ndarray) else frames)
        t = t.to(eng.device).contiguous()
        T = t.shape[0]
        det_logits = eng.forward(t)
        if reset:
            eng.convlstm_reset()
        trk_logits = eng.convlstm_window(T)
        out = []
        for lg in (trk_logits, det_logits):
            boxes, counts = eng.decode(lg, self.OBJ_THRESHOLD, self.NMS_THRESHOLD, self.ANCHORS)
            out.append([boxes_from_rows(r, self.CLASS) for r in rows_to_host(boxes, counts)])
        return out[0], out[1]

    def predict(self, input_paths, output_paths):
        """MultiObjDetTracker.py:295-315 (with its evident intent: one window of SEQUENCE_LENGTH frames)."""
        assert len(input_paths) == self.SEQUENCE_LENGTH and len(output_paths) == self.SEQUENCE_LENGTH
        images = [load_frame(p) for p in input_paths]
        # cv2.resize (MultiObjDetTracker.py:300) on the device, bit-identical to OpenCV's INTER_LINEAR
        dev = self.model.device
        x = torch.cat([self.model.resize_frames(torch.from_numpy(np.ascontiguousarray(im[None])).to(dev), self.IMAGE_H)
                       if im.shape[:2] != (self.IMAGE_H, self.IMAGE_H) else torch.from_numpy(np.ascontiguousarray(im[None])).to(dev)
                       for im in images])
        trk, _ = self.track_window(x)
        import cv2
        for image, boxes, path in zip(images, trk, output_paths):
            image = draw_boxes(image, boxes, self.LABELS)
            print(len(boxes), 'Bounding Boxes Found')
            print("File Saved to", path)
            cv2.imwrite(path, image)
        return trk
