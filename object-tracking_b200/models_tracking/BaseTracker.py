"""Drop-in for the inference-relevant half of ``models_tracking/BaseTracker.py`` (:12-60) plus the per-frame
``step()/reset()`` the reference lacks: the dataflow of
``BatchSequenceGenerator2.output_from_instance`` (utility/preprocessing.py:403-477) -- detect, take the
highest-probability detection of an allowed class, normalise it by the frame size, pool the fv_layer
feature, feed the recurrent head -- without the JPEG round-trip through disk (:412-415).

Everything numeric runs on the GPU: detector forward, region decode + NMS, detection choice, feature
pooling, LSTM cell, Dense head.  ``train()`` is out of scope.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from ..models_detection.YOLO import YOLO
from ..models_detection._common import load_config


class BaseTracker(object):
    def __init__(self, config=None, max_streams: int = 1, detector_kwargs: Optional[dict] = None,
                 ref_layout_bug: bool = False):
        self.config = load_config(config)
        self.detection_model = self.config["model_detector"]["name"]
        self.detection_fv_layer = self.config["model_detector"]["fv_layer"]
        self.cpu_mode = self.config["train"]["cpu_only"]
        self.tgpu_id = self.config["train"]["tgpu_id"]
        self.dgpu_id = self.config["train"]["dgpu_id"]
        self.pool = self.config["train"]["pool"]
        self.batch_size = self.config["train"]["batch_size"]
        self.max_epochs = self.config["train"]["max_epochs"]
        self.sequence_length = self.config["model_tracker"]["sequence_length"]
        self.classes = self.config["train"]["classes"]
        self.model_name = self.config["model_tracker"]["name"]
        self.tensorboard_dir = self.config["train"]["tensorboard_dir"]
        self.saved_model_path = self.config["train"]["saved_model_dir"] + self.model_name
        self.max_streams = max_streams
        self.ref_layout_bug = ref_layout_bug          # preprocessing.py:419 CHW-as-HWC reinterpretation (R11)
        self._detector_kwargs = dict(detector_kwargs or {})
        self._graphs = {}
        self.load_detection_model()

    def load_detection_model(self):
        if self.detection_model != 'YOLO':
            raise NotImplementedError("only the YOLO detector plugin is built (FasterRCNN: SURVEY.md 8f rank 4)")
        kw = dict(self._detector_kwargs)
        kw.setdefault("max_batch", self.max_streams * self.sequence_length)
        self.model_detector = YOLO([self.cpu_mode, self.dgpu_id], config=self.config, **kw)
        self._w, self._h, self._c = self.model_detector.get_layer_dims(self.detection_fv_layer)
        self._fv_name = self.model_detector._name(self.detection_fv_layer)

    def load_tracker_model(self):
        raise NotImplementedError

    def load_weights(self, path: Optional[str] = None):
        """Load the LSTM + Dense head from a checkpoint: the Keras ``.hdf5`` files ``train()`` writes
        (BaseTracker.py:74-80, ``<saved_model_dir><name>-CHKPNT-<epoch>-<val_loss>.hdf5``; read without h5py) or an
        ``.npz``.  path=None takes the latest checkpoint under ``self.saved_model_path``.  Returns the file used."""
        from ..weights import latest_checkpoint, load_checkpoint_arrays, lstm_weights_from_arrays
        path = path or latest_checkpoint(self.saved_model_path)
        if path is None:
            raise FileNotFoundError(f"no checkpoint {self.saved_model_path}-CHKPNT-*.{{hdf5,h5,npz}}")
        self.head.set_weights(lstm_weights_from_arrays(load_checkpoint_arrays(path)))
        self.head.reset(-1)
        return path

    def load_data_generators(self):
        raise NotImplementedError("training data generators are out of scope of the B200 hot path")

    def train(self):
        raise NotImplementedError("training is out of scope of the B200 hot path (SURVEY.md section 2.1)")

    # ------------------------------------------------------------------ feature size of the pooled fv
    def _n_feat(self) -> int:
        if self.pool == 'Global':
            return self._c
        if self.pool == 'Max':
            return (self._w // 4) * (self._h // 4) * self._c
        raise ValueError(f"unknown pool '{self.pool}'")

    # ------------------------------------------------------------------ new per-frame API
    def reset(self, stream: int = -1):
        self.head.reset(stream)
        steps = self.__dict__.setdefault("_stream_steps", {})
        if stream < 0:
            steps.clear()
        else:
            steps[stream] = 0

    def _detect_and_pool(self, frames: torch.Tensor, heat_size: int = 0):
        self.model_detector.engine.forward(frames)
        return self._decode_and_pool(frames.shape[0], frames.shape[2], frames.shape[1], heat_size)

    def _decode_and_pool(self, B: int, W: int, H: int, heat_size: int = 0):
        """Everything after the conv stack, on the engine's current logits / feature buffers."""
        det = self.model_detector
        # the feature pooling does not depend on the detections: it runs on a side stream beside decode + selection
        # (two parallel branches when the tail is captured into a CUDA graph)
        cur = torch.cuda.current_stream()
        side = self.__dict__.get("_pool_stream")
        if side is None:
            side = self._pool_stream = torch.cuda.Stream(device=det.engine.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            fv = det.engine.pool_features(self._fv_name, B, self.pool, self.ref_layout_bug)
        dets, counts = det.engine.region_detect(det.engine.logits(B), det.THRESH, det.NMS, W, H)
        det_in, heat, chosen = det.engine.select_detection(dets, counts, W, H, det.class_mask, heat_size)
        cur.wait_stream(side)
        return fv, det_in, heat, chosen

    def _tail(self, S: int, T: int, W: int, H: int, reset: bool) -> torch.Tensor:
        fv, xin = self._tracker_inputs_from_state(S * T, W, H)
        return self.head.sequence(fv.view(S, T, -1), xin.view(S, T, -1), reset=reset)

    def track_windows(self, frames: torch.Tensor, reset: bool = True, graph: bool = True,
                      pipeline: bool = False) -> torch.Tensor:
        """frames (S,T,H,W,3) uint8 on the GPU: S independent streams (or windows), T consecutive frames each.
        One batched detector pass over S*T frames, then T recurrent steps over the S streams in parallel
        (input projection hoisted out of the recurrence).  reset=True reproduces Keras' stateless windows.
        graph=True replays the step from two CUDA graphs captured on first use (conv stack | decode + tracker):
        the ~45 launches of a step are launch-latency bound otherwise.  The result tensor is reused.
        pipeline=True (with graph=True) additionally overlaps the tracker tail of this call with the first conv
        layers of the NEXT call: the tail (decode, selection, pooling, LSTM) runs on ``self.tail_stream`` and only
        reads outputs of conv_9 and later, so the next call's conv_1..8 start at once and its conv_9..23 wait for
        ``self.tail_done``.  The returned tensor is then valid after ``self.tail_done`` (an event on the tail stream):
        consume it on ``self.tail_stream`` or wait for the event."""
        S, T, H, W = frames.shape[0], frames.shape[1], frames.shape[2], frames.shape[3]
        if S > self.max_streams:
            raise ValueError(f"{S} streams > max_streams {self.max_streams}")
        eng = self.model_detector.engine
        if not graph:
            eng.forward(frames.reshape(S * T, H, W, 3))
            return self._tail(S, T, W, H, reset)
        if pipeline:
            return self._track_windows_pipelined(frames, reset)
        if getattr(self, "tail_done", None) is not None:       # a pipelined call may still be in its tail
            torch.cuda.current_stream().wait_event(self.tail_done)
            self.tail_done = None
        # uint8 windows whose per-stream blocks are contiguous are gathered by the ingest kernel itself (no staging
        # copy; the graph starts at conv_1); anything else is copied into a static input tensor first
        ingest = eng.can_ingest(frames)
        key = (S, T, H, W, frames.dtype, bool(reset), ingest)
        g = self._graphs.get(key)
        if g is None:
            static_in = None
            if not ingest:
                static_in = torch.empty((S * T, H, W, 3), dtype=frames.dtype, device=frames.device)
                static_in.view(S, T, H, W, 3).copy_(frames)
            fwd = (lambda: eng.forward_ingested(S * T)) if ingest else (lambda: eng.forward(static_in))
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                      # warm-up outside capture (allocations, lazy init)
                if ingest:
                    eng.ingest_windows(frames)
                fwd()
                self._tail(S, T, W, H, reset)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            ev, eng.forward_events = eng.forward_events, None   # no event records inside a capture
            n0 = eng.lib.b2t_launch_count(eng.h)
            g_fwd, g_tail = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_fwd):
                fwd()
            with torch.cuda.graph(g_tail):
                static_out = self._tail(S, T, W, H, reset)
            eng.forward_events = ev
            g = self._graphs[key] = (g_fwd, g_tail, static_in, static_out, eng.lib.b2t_launch_count(eng.h) - n0)
        g_fwd, g_tail, static_in, static_out, n_kernels = g
        ev = eng.forward_events
        if ev is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if ingest:
            eng.ingest_windows(frames)
        else:
            static_in.view(S, T, H, W, 3).copy_(frames, non_blocking=True)
        g_fwd.replay()
        if ev is not None:
            e1.record()
            ev.append((e0, e1))
        g_tail.replay()
        eng.add_graph_launches(n_kernels)
        return static_out

    _SPLIT = 8          # conv_1 .. conv_8 do not write anything the tracker tail reads

    def _track_windows_pipelined(self, frames: torch.Tensor, reset: bool) -> torch.Tensor:
        S, T, H, W = frames.shape[0], frames.shape[1], frames.shape[2], frames.shape[3]
        eng = self.model_detector.engine
        ingest = eng.can_ingest(frames)
        key = ("pipe", S, T, H, W, frames.dtype, bool(reset), ingest)
        g = self._graphs.get(key)
        if g is None:
            static_in = None
            if not ingest:
                static_in = torch.empty((S * T, H, W, 3), dtype=frames.dtype, device=frames.device)
                static_in.view(S, T, H, W, 3).copy_(frames)
            if ingest:
                fwd_a = lambda: eng.forward_ingested(S * T, 1, self._SPLIT)
                fwd_b = lambda: eng.forward_ingested(S * T, self._SPLIT + 1, 23)
            else:
                fwd_a = lambda: eng.forward_range(static_in, 1, self._SPLIT)
                fwd_b = lambda: eng.forward_range(static_in, self._SPLIT + 1, 23)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                      # warm-up outside capture (allocations, lazy init)
                if ingest:
                    eng.ingest_windows(frames)
                fwd_a()
                fwd_b()
                self._tail(S, T, W, H, reset)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            ev, eng.forward_events = eng.forward_events, None
            n0 = eng.lib.b2t_launch_count(eng.h)
            g_a, g_b, g_tail = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_a):
                fwd_a()
            with torch.cuda.graph(g_b):
                fwd_b()
            with torch.cuda.graph(g_tail):
                static_out = self._tail(S, T, W, H, reset)
            eng.forward_events = ev
            g = self._graphs[key] = (g_a, g_b, g_tail, static_in, static_out, eng.lib.b2t_launch_count(eng.h) - n0)
            if getattr(self, "tail_stream", None) is None:
                self.tail_stream = torch.cuda.Stream()
                self.tail_done = None
        g_a, g_b, g_tail, static_in, static_out, n_kernels = g
        cur = torch.cuda.current_stream()
        ev = eng.forward_events
        if ev is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if ingest:
            eng.ingest_windows(frames)
        else:
            static_in.view(S, T, H, W, 3).copy_(frames, non_blocking=True)
        g_a.replay()                                           # conv_1 .. conv_8: overlaps the previous call's tail
        if self.tail_done is not None:
            cur.wait_event(self.tail_done)                     # conv_9 .. 23 overwrite what that tail reads
        g_b.replay()
        if ev is not None:
            e1.record()
            ev.append((e0, e1))
        conv_done = torch.cuda.Event()
        conv_done.record(cur)
        with torch.cuda.stream(self.tail_stream):
            self.tail_stream.wait_event(conv_done)
            g_tail.replay()
            self.tail_done = torch.cuda.Event()
            self.tail_done.record(self.tail_stream)
        eng.add_graph_launches(n_kernels)
        return static_out

    def step(self, frame, det_bbox=None, stream: int = 0) -> np.ndarray:
        """One frame of stream `stream` (online use).  frame: HWC uint8 array or GPU tensor.
        det_bbox: optional [cx, cy, w, h] in pixels of the frame that REPLACES the detector's own choice (e.g. the
        ground-truth box of the first frame, or an external detector's output); None = the highest-probability
        detection of an allowed class (zeros if none, preprocessing.py:434-449).
        The (h, c) state of every stream persists in its own slot of the head (max_streams slots) and is reset every
        ``sequence_length`` steps of that stream, like the reference's stateless 4-frame windows."""
        if not 0 <= stream < self.max_streams:
            raise ValueError(f"stream {stream} outside [0, max_streams={self.max_streams})")
        eng = self.model_detector.engine
        t = torch.as_tensor(np.ascontiguousarray(frame) if isinstance(frame, np.ndarray) else frame)
        t = t.to(eng.device)
        steps = self.__dict__.setdefault("_stream_steps", {})
        n = steps.get(stream, 0)
        if n % self.sequence_length == 0:
            self.head.reset(stream)
        steps[stream] = n + 1
        t = t[None].contiguous()
        H, W = t.shape[1], t.shape[2]
        eng.forward(t)
        fv, xin = self._tracker_inputs_from_state(1, W, H)
        if det_bbox is not None:
            box = torch.tensor([[float(det_bbox[0]) / W, float(det_bbox[1]) / H, float(det_bbox[2]) / W,
                                 float(det_bbox[3]) / H]], dtype=torch.float32, device=eng.device)
            if xin.shape[1] == 4:
                xin = box
            else:                                   # heat-map trackers take the box's top-left corner (:452-456)
                tl = torch.cat([box[:, :2] - box[:, 2:] / 2, box[:, 2:]], dim=1).contiguous()
                xin = eng.heatmap_from_box(tl, int(round(xin.shape[1] ** 0.5)))
        return self.head.step(fv, xin, slot0=stream)[0].cpu().numpy()
