"""Drop-in for the inference half of ``models_tracking/MultiObjDetTracker.py``: TimeDistributed YOLOv2 ->
concat[conv_23 logits, conv_feat] -> ConvLSTM2D(512, 3x3, same, return_sequences) -> Conv2D(5*(5+C), 1x1)
-> Reshape -> decode_netout per frame (MultiObjDetTracker.py:160-189, :295-315).  Losses / fit loop are out
of scope.  The ConvLSTM and the head are the same tcgen05 implicit-GEMM kernel as the backbone.

Kept surface: class constants (:82-106), ``MultiObjDetTracker(argv={})``, ``.detector`` (a KerasYOLO built with
BATCH_SIZE*SEQUENCE_LENGTH), ``.model_detector``, ``.model``, ``.predict(input_paths, output_paths)``.
Added: ``track_windows`` (S streams x T frames per call, one CUDA graph per geometry),
``step(frame, stream)`` (online, state persists per stream, reset every SEQUENCE_LENGTH steps), ``reset``."""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from ..engine import rows_to_host
from ..models_detection.KerasYOLO import KerasYOLO
from ..models_detection._common import load_frame
from ..utility.utils import BoundBox, boxes_from_rows, draw_boxes
from ..weights import ANCHORS, synthetic_multiobj_weights


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
    SAVED_MODEL_PATH = 'models/MultiObjDetTracker-CHKPNT-03-0.55.hdf5'
    model = None
    detector = None
    model_detector = None

    def __init__(self, argv={}, device: int = 0, detector_weights: Optional[dict] = None,
                 tracker_weights: Optional[dict] = None, convlstm_units: Optional[int] = None,
                 max_streams: Optional[int] = None):
        if 'LABELS' in argv:
            self.LABELS = argv['LABELS']
        self.CLASS = len(self.LABELS)
        if convlstm_units:
            self.CONVLSTM_UNITS = convlstm_units
        if max_streams:                       # streams (or independent windows) batched per call = Keras' BATCH_SIZE
            self.BATCH_SIZE = max_streams
        argv = dict(argv)
        argv['LABELS'] = self.LABELS
        argv['BATCH_SIZE'] = self.BATCH_SIZE * self.SEQUENCE_LENGTH
        argv['IMAGE_H'], argv['IMAGE_W'] = self.IMAGE_H, self.IMAGE_W
        argv['GRID_H'], argv['GRID_W'] = self.GRID_H, self.GRID_W
        self.detector = KerasYOLO(argv, device=device, convlstm_units=self.CONVLSTM_UNITS, weights=detector_weights)
        self._tracker_weights = tracker_weights
        self._graphs = {}
        self._steps = {}
        self.load_model()

    def load_model(self):
        eng = self.detector.model
        w = self._tracker_weights
        if w is None and self.LOAD_MODEL:                  # MultiObjDetTracker.py:132-133
            w = self._checkpoint_weights(self.SAVED_MODEL_PATH)
        if w is None:
            w = synthetic_multiobj_weights(self.CLASS, self.CONVLSTM_UNITS, seed=2)
        eng.set_convlstm_weights(w)
        eng.finalize()
        self.model = self.model_detector = eng
        eng.convlstm_reset()

    def _checkpoint_weights(self, path: str):
        """MultiObjDetTracker.py:291-293: the checkpoint holds the whole model -- ConvLSTM2D + head and, when present,
        the detector's conv_k / norm_k layers (set before finalize)."""
        from ..weights import convlstm_weights_from_arrays, detector_weights_from_arrays, load_checkpoint_arrays
        arrays = load_checkpoint_arrays(path)
        det = detector_weights_from_arrays(arrays, self.CLASS)
        if det is not None:
            self.model.set_weights(det) if self.model is not None else self.detector.model.set_weights(det)
        import re
        m = re.search(r"-CHKPNT-(\d+)-", path)                # MultiObjDetTracker.py:293: int(path.split('-')[2])
        self.INITIAL_EPOCH = int(m.group(1)) if m else 0
        return convlstm_weights_from_arrays(arrays)

    def load_weights(self, path: Optional[str] = None):
        """Keras ``.hdf5`` (read without h5py) or ``.npz`` checkpoint -> ConvLSTM2D + head (+ detector layers if the
        file has them); re-packs the weight blob."""
        w = self._checkpoint_weights(path or self.SAVED_MODEL_PATH)
        self.model.set_convlstm_weights(w)
        self.model.finalize()
        self.model.convlstm_reset()
        self._graphs = {}

    def reset(self, stream: int = -1):
        if stream < 0:
            self.model.convlstm_reset()
            self._steps = {}
        else:
            self.model.convlstm_reset_slots(stream, 1)
            self._steps[stream] = 0

    # ------------------------------------------------------------------ batched device API
    def _window_kernels(self, static_in: Optional[torch.Tensor], S: int, T: int, reset: bool, decode_detector: bool):
        """static_in None: the frames were placed by eng.ingest_windows()."""
        eng = self.model
        det_logits = eng.forward(static_in) if static_in is not None else eng.forward_ingested(S * T)
        trk_logits = eng.convlstm_sequence(S, T, 0, reset)
        boxes, counts = eng.decode(trk_logits, self.OBJ_THRESHOLD, self.NMS_THRESHOLD, self.ANCHORS, tag="trk")
        out = [trk_logits, boxes, counts]
        if decode_detector:
            b2, c2 = eng.decode(det_logits, self.OBJ_THRESHOLD, self.NMS_THRESHOLD, self.ANCHORS, tag="det")
            out += [b2, c2]
        return out

    def track_windows(self, frames: torch.Tensor, reset: bool = True, graph: bool = True):
        """frames (S,T,H,W,3) uint8 on the GPU: S independent streams (or windows) of T consecutive frames.
        -> (trk_logits (S*T,G,G,A,5+C), boxes (S*T,max,8), counts (S*T)) device tensors, frame index s*T + t:
        the tracker output decoded like MultiObjDetTracker.predict (:309-310).  One batched detector pass, one
        batched ConvLSTM input conv, T recurrent steps over the S streams, one batched head, one decode launch;
        with graph=True the whole call replays as ONE CUDA graph captured on first use (uint8 windows are gathered by
        the ingest kernel in front of it, no staging copy).  The tensors are reused."""
        S, T, H, W = frames.shape[0], frames.shape[1], frames.shape[2], frames.shape[3]
        if S * T > self.detector.BATCH_SIZE:
            raise ValueError(f"{S} streams x {T} frames > detector batch {self.detector.BATCH_SIZE}")
        eng = self.model
        if not graph:
            return tuple(self._window_kernels(frames.reshape(S * T, H, W, 3).contiguous(), S, T, reset, False))
        ingest = eng.can_ingest(frames)
        key = (S, T, H, W, frames.dtype, bool(reset), ingest)
        g = self._graphs.get(key)
        if g is None:
            static_in = None
            if not ingest:
                static_in = torch.empty((S * T, H, W, 3), dtype=frames.dtype, device=frames.device)
                static_in.view(S, T, H, W, 3).copy_(frames)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                      # warm-up outside capture (allocations, lazy init)
                if ingest:
                    eng.ingest_windows(frames)
                self._window_kernels(static_in, S, T, reset, False)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            ev, eng.forward_events = eng.forward_events, None   # no event records inside a capture
            n0 = eng.lib.b2t_launch_count(eng.h)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                outs = self._window_kernels(static_in, S, T, reset, False)
            eng.forward_events = ev
            g = self._graphs[key] = (gr, static_in, outs, eng.lib.b2t_launch_count(eng.h) - n0)
        gr, static_in, outs, n_kernels = g
        ev = eng.forward_events
        if ev is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if ingest:
            eng.ingest_windows(frames)
        else:
            static_in.view(S, T, H, W, 3).copy_(frames, non_blocking=True)
        gr.replay()
        if ev is not None:
            e1.record()
            ev.append((e0, e1))
        eng.add_graph_launches(n_kernels)
        return tuple(outs)

    def track_window(self, frames, reset: bool = True):
        """frames (T,H,W,3) uint8 (T <= SEQUENCE_LENGTH): -> (tracker boxes per frame, detector boxes per frame)."""
        eng = self.model
        t = torch.as_tensor(np.ascontiguousarray(frames) if isinstance(frames, np.ndarray) else frames)
        t = t.to(eng.device).contiguous()
        T = t.shape[0]
        _, boxes, counts, b2, c2 = self._window_kernels(t, 1, T, reset, True)
        trk = [boxes_from_rows(r, self.CLASS) for r in rows_to_host(boxes, counts)]
        det = [boxes_from_rows(r, self.CLASS) for r in rows_to_host(b2, c2)]
        return trk, det

    # ------------------------------------------------------------------ online API (absent in the reference)
    def step(self, frame, stream: int = 0) -> List[BoundBox]:
        """One frame of stream `stream` (HWC uint8 array or GPU tensor) -> the tracker's boxes for that frame.  The
        ConvLSTM state of the stream persists between calls and is reset every SEQUENCE_LENGTH steps, which
        reproduces the reference's stateless SEQUENCE_LENGTH-frame windows."""
        eng = self.model
        if not 0 <= stream < self.detector.BATCH_SIZE:
            raise ValueError(f"stream {stream} outside [0, {self.detector.BATCH_SIZE})")
        t = torch.as_tensor(np.ascontiguousarray(frame) if isinstance(frame, np.ndarray) else frame)
        t = t.to(eng.device)[None].contiguous()
        n = self._steps.get(stream, 0)
        eng.forward(t)
        trk_logits = eng.convlstm_sequence(1, 1, stream, reset=(n % self.SEQUENCE_LENGTH == 0))
        self._steps[stream] = n + 1
        boxes, counts = eng.decode(trk_logits, self.OBJ_THRESHOLD, self.NMS_THRESHOLD, self.ANCHORS)
        return boxes_from_rows(rows_to_host(boxes, counts)[0], self.CLASS)

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
