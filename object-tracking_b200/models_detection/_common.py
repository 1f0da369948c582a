"""Shared host glue of the detector plugins: config lookup, label lists, frame loading."""
from __future__ import annotations

import json
import os
from typing import Optional, Union

import numpy as np

COCO_NAMES = [
    "person", "bicycle", "car", "motorbike", "aeroplane", "bus", "train", "truck", "boat", "traffic light",
    "fire hydrant", "stop sign", "parking meter", "bench", "bird", "cat", "dog", "horse", "sheep", "cow",
    "elephant", "bear", "zebra", "giraffe", "backpack", "umbrella", "handbag", "tie", "suitcase", "frisbee",
    "skis", "snowboard", "sports ball", "kite", "baseball bat", "baseball glove", "skateboard", "surfboard",
    "tennis racket", "bottle", "wine glass", "cup", "fork", "knife", "spoon", "bowl", "banana", "apple",
    "sandwich", "orange", "broccoli", "carrot", "hot dog", "pizza", "donut", "cake", "chair", "sofa",
    "pottedplant", "bed", "diningtable", "toilet", "tvmonitor", "laptop", "mouse", "remote", "keyboard",
    "cell phone", "microwave", "oven", "toaster", "sink", "refrigerator", "book", "clock", "vase", "scissors",
    "teddy bear", "hair drier", "toothbrush"]           # darknet/data/coco.names order

DEFAULT_CONFIG = {
    "model_detector": {"name": "YOLO", "config_file": "cfg/yolov2.cfg", "meta_file": "cfg/coco.data",
                       "weights_file": "yolov2.weights", "fv_layer": 25, "nms": 0.45, "thresh": 0.5,
                       "hier_thresh": 0.5},
    "model_tracker": {"name": "TinyTracker", "lstm_units": 512, "sequence_length": 4, "heatmap_size": 32},
    "train": {"cpu_only": 0, "dgpu_id": 0, "tgpu_id": 0, "pool": "Global", "batch_size": 4, "max_epochs": 100,
              "tensorboard_dir": "logs/", "saved_model_dir": "models/", "classes": ["Person", "Car"], "debug": 0,
              "train_image_folder": "data/VisualTB/", "train_annot_folder": "data/VisualTBAnn/train/"},
    "val": {"val_image_folder": "data/VisualTB/", "val_annot_folder": "data/VisualTBAnn/val/"},
}


def load_config(config: Union[None, str, dict] = None) -> dict:
    """The reference opens ``config.json`` in the working directory (BaseTracker.py:14, YOLO.py:41).  Same
    here; a dict or an explicit path may be passed instead, and the reference's stock values are the
    fallback when no file exists (there is no config.json on a bench box)."""
    if isinstance(config, dict):
        return config
    path = config or os.environ.get("B2T_CONFIG", "config.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.loads(f.read())
    return json.loads(json.dumps(DEFAULT_CONFIG))


def load_frame(frame, size: Optional[int] = None) -> np.ndarray:
    """Path -> HWC uint8 via OpenCV (BGR, as cv2.imread gives it); arrays pass through.  The hot path takes
    arrays/tensors; file decoding is the reference's own host step (KerasYOLO.py:525-526)."""
    if isinstance(frame, str):
        import cv2
        img = cv2.imread(frame)
        if img is None:
            raise IOError(f"cannot read image {frame}")
        frame = img
    frame = np.asarray(frame)
    if size is not None and frame.shape[:2] != (size, size):
        import cv2
        frame = cv2.resize(frame, (size, size))
    return frame
