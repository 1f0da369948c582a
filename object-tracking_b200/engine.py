"""Thin host wrapper over libb200track.so: torch tensors are the device-memory container and the
stream provider, every numeric step is a call through the C-ABI (``_native``).

``DetectorEngine`` = one context per GPU: weight blob + workspace (two torch uint8 tensors), the YOLOv2
forward, decode/NMS (keras and darknet flavours), feature pooling, ConvLSTM window.
``LstmHead``       = the LSTM + Dense head of TinyTracker / TinyHeatmapTracker with persistent (h, c).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as N
from .weights import ANCHORS, N_BOX, yolo_layer_table


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _aligned_bytes(n: int, device) -> torch.Tensor:
    """uint8 tensor of n bytes whose data_ptr is 1024-byte aligned (TMA/swizzle atoms)."""
    raw = torch.empty(n + 1024, dtype=torch.uint8, device=device)
    off = (-raw.data_ptr()) % 1024
    t = raw[off:off + n]
    t._b2t_keepalive = raw
    return t


def _np_ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def rows_to_host(rows: torch.Tensor, counts: torch.Tensor):
    """(B,max,8) decode rows + (B,) counts on the device -> list of per-frame (n,8) numpy arrays.  The decode kernels
    report a candidate / entry overflow of their shared-memory tables as count = -1 (only reachable with thresholds
    below 0.5, where several classes per anchor can pass): that is an error, never a slice."""
    n = counts.cpu().numpy()
    if (n < 0).any():
        raise N.B2TError("decode: more (anchor, class) candidates than the kernel's capacity in frame(s) "
                         f"{np.nonzero(n < 0)[0].tolist()} (threshold too low)")
    r = rows.cpu().numpy()
    return [r[i, :int(n[i])] for i in range(len(n))]


class DetectorEngine:
    def __init__(self, n_class: int = 80, image_size: int = 416, max_batch: int = 4, semantics: str = "keras",
                 bn_eps: float = 1e-3, engine: str = "tcgen05", device: int = 0, convlstm_units: int = 0,
                 keep_prepool: bool = False, chain_max_batch: int = 0, graph: str = "yolov2", tiny_filters: int = 1024):
        """graph: "yolov2" (cfg/yolov2.cfg, yolov2-voc.cfg: 23 conv layers) or "tiny" (cfg/yolov2-tiny*.cfg: 9 conv
        layers; tiny_filters = filters of its conv_8, 1024 voc / 512 coco; weights through load_darknet_weights).
        chain_max_batch: largest batch whose conv_2..23 run as one persistent cooperative launch (the small-batch
        schedule, conv_chain_kernel); 0 = library default (1), -1 = never."""
        if not torch.cuda.is_available():
            raise N.B2TError("no CUDA device: the B200 path has no CPU fallback")
        self.lib = N.lib()
        self.n_class, self.image_size, self.max_batch = n_class, image_size, max_batch
        self.grid = image_size // 32
        self.n_box = N_BOX
        self.device = torch.device("cuda", device)
        cfg = N.Config()
        cfg.image_h = cfg.image_w = image_size
        cfg.n_class, cfg.max_batch = n_class, max_batch
        cfg.semantics = {"keras": N.SEM_KERAS, "darknet": N.SEM_DARKNET}[semantics]
        cfg.bn_eps = bn_eps
        cfg.engine = {"tcgen05": N.ENGINE_TCGEN05, "simt": N.ENGINE_SIMT, "tcgen05_tile": N.ENGINE_TCGEN05_TILE}[engine]
        cfg.device = device
        cfg.convlstm_units = convlstm_units
        cfg.reserved[0] = 1 if keep_prepool else 0
        cfg.reserved[1] = chain_max_batch
        if graph not in ("yolov2", "tiny"):
            raise ValueError(f"unknown graph '{graph}'")
        if graph == "tiny":
            cfg.reserved[2], cfg.reserved[3] = 1, tiny_filters
        self.graph = graph
        self.semantics, self.convlstm_units = semantics, convlstm_units
        h = C.c_void_p()
        N.check(self.lib.b2t_create(C.byref(cfg), C.byref(h)))
        self.h = h
        torch.cuda.set_device(self.device)
        self.blob = _aligned_bytes(self.lib.b2t_weight_bytes(h), self.device)
        self.workspace = _aligned_bytes(self.lib.b2t_workspace_bytes(h), self.device)
        N.check(self.lib.b2t_bind_memory(h, self.blob.data_ptr(), self.workspace.data_ptr()))
        self.finalized = False
        self._anchors = (C.c_float * (2 * N_BOX))(*ANCHORS)
        self._scratch: Dict[Tuple, torch.Tensor] = {}
        self._logits_view = None
        self.forward_events = None
        self._graph_launches = 0

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.b2t_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---------------------------------------------------------------- weights
    def set_weights(self, w: Dict[str, np.ndarray]) -> None:
        """Keras-layout arrays (weights.synthetic_yolo_weights / read_darknet_weights)."""
        for s in yolo_layer_table(self.n_class):
            k = np.ascontiguousarray(w[f"kernel_{s.index}"], dtype=np.float32)
            if k.shape != (s.ksize, s.ksize, s.cin, s.cout):
                raise ValueError(f"kernel_{s.index}: shape {k.shape} != {(s.ksize, s.ksize, s.cin, s.cout)}")
            if s.bn:
                arrs = [np.ascontiguousarray(w[f"{n}_{s.index}"], dtype=np.float32) for n in ("gamma", "beta", "mean", "var")]
                N.check(self.lib.b2t_set_conv_weights(self.h, s.index, _np_ptr(k), *[_np_ptr(a) for a in arrs], None))
            else:
                b = np.ascontiguousarray(w[f"bias_{s.index}"], dtype=np.float32)
                N.check(self.lib.b2t_set_conv_weights(self.h, s.index, _np_ptr(k), None, None, None, None, _np_ptr(b)))

    def load_darknet_weights(self, path: str) -> None:
        N.check(self.lib.b2t_load_darknet_weights(self.h, path.encode()))

    def set_convlstm_weights(self, w: Dict[str, np.ndarray]) -> None:
        a = [np.ascontiguousarray(w[k], dtype=np.float32) for k in
             ("kernel", "recurrent_kernel", "bias", "head_kernel", "head_bias")]
        N.check(self.lib.b2t_set_convlstm_weights(self.h, *[_np_ptr(x) for x in a]))

    def finalize(self, upload: bool = True) -> None:
        """Upload the packed blob (rank 0) or keep what a broadcast wrote (other ranks); build TMA maps."""
        N.check(self.lib.b2t_finalize(self.h, 1 if upload else 0, _stream()))
        torch.cuda.current_stream().synchronize()
        self.finalized = True

    def broadcast_weights(self, src: int = 0) -> None:
        """The only collective on the path: one NCCL broadcast of the packed blob at init."""
        from .sharding import broadcast_blob
        broadcast_blob(self.blob, src=src)

    # ---------------------------------------------------------------- forward
    def forward(self, frames: torch.Tensor) -> torch.Tensor:
        """frames (B,H,W,3) uint8 or float32 on this GPU -> logits (B,G,G,A,5+C) fp32 (engine-owned view)."""
        if frames.device != self.device or not frames.is_contiguous():
            raise ValueError("frames must be a contiguous tensor on the engine's device")
        if frames.dim() != 4 or tuple(frames.shape[1:]) != (self.image_size, self.image_size, 3):
            raise ValueError(f"frames shape {tuple(frames.shape)} != (B,{self.image_size},{self.image_size},3)")
        dt = {torch.uint8: N.FRAME_U8, torch.float32: N.FRAME_F32}.get(frames.dtype)
        if dt is None:
            raise ValueError("frames must be uint8 or float32")
        B = frames.shape[0]
        ev = self.forward_events
        if ev is not None:                      # bench.py: device time of the conv stack, per call
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        N.check(self.lib.b2t_yolo_forward(self.h, frames.data_ptr(), dt, B, None, _stream()))
        if ev is not None:
            b.record()
            ev.append((a, b))
        return self.logits(B)

    def forward_range(self, frames: torch.Tensor, first: int, last: int) -> None:
        """conv_first .. conv_last of the forward pass over `frames` (same checks as forward())."""
        if frames.device != self.device or not frames.is_contiguous():
            raise ValueError("frames must be a contiguous tensor on the engine's device")
        dt = {torch.uint8: N.FRAME_U8, torch.float32: N.FRAME_F32}.get(frames.dtype)
        if dt is None:
            raise ValueError("frames must be uint8 or float32")
        N.check(self.lib.b2t_yolo_forward_range(self.h, frames.data_ptr(), dt, frames.shape[0], first, last, None, _stream()))

    # ---------------------------------------------------------------- ingest without a staging copy
    def can_ingest(self, frames: torch.Tensor) -> bool:
        """(S,T,H,W,3) uint8 windows whose per-stream blocks frames[s] are contiguous (views of longer clips are)."""
        if frames.dim() != 5 or frames.dtype != torch.uint8 or frames.device != self.device:
            return False
        S, T, H, W, C3 = frames.shape
        if (H, W, C3) != (self.image_size, self.image_size, 3) or S * T > self.max_batch:
            return False
        st = frames.stride()
        return st[4] == 1 and st[3] == 3 and st[2] == 3 * W and st[1] == 3 * W * H and (S == 1 or st[0] >= T * H * W * 3)

    def ingest_windows(self, frames: torch.Tensor) -> None:
        """Gather S windows of T frames straight into the engine's input buffer (no staging copy); follow with
        forward_ingested(S*T)."""
        S, T = frames.shape[0], frames.shape[1]
        N.check(self.lib.b2t_ingest_frames(self.h, frames.data_ptr(), S, T, frames.stride(0) if S > 1 else 0, _stream()))

    def forward_ingested(self, B: int, first: int = 1, last: int = 23) -> torch.Tensor:
        """conv_first..conv_last on the frames ingest_windows() placed in the input buffer.  Capturable in a CUDA graph
        whose input changes every replay."""
        ev = self.forward_events
        if ev is not None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        N.check(self.lib.b2t_yolo_forward_range(self.h, None, N.FRAME_U8, B, first, last, None, _stream()))
        if ev is not None:
            b.record()
            ev.append((a, b))
        return self.logits(B)

    def logits(self, B: int) -> torch.Tensor:
        """Zero-copy view of the context's logits buffer (it lives inside the torch-owned workspace)."""
        if self._logits_view is None:
            off = self.lib.b2t_logits(self.h) - self.workspace.data_ptr()
            n = self.max_batch * self.grid * self.grid * self.n_box * (5 + self.n_class)
            self._logits_view = self.workspace[off:off + 4 * n].view(torch.float32).view(
                self.max_batch, self.grid, self.grid, self.n_box, 5 + self.n_class)
        return self._logits_view[:B]

    def _buf(self, key, shape, dtype) -> torch.Tensor:
        t = self._scratch.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._scratch[key] = t
        return t

    def layer_dims(self, name: str) -> Tuple[int, int, int]:
        h, w, c = C.c_int(), C.c_int(), C.c_int()
        N.check(self.lib.b2t_layer_dims(self.h, name.encode(), C.byref(h), C.byref(w), C.byref(c)))
        return h.value, w.value, c.value

    def extract(self, name: str, B: int) -> torch.Tensor:
        h, w, c = self.layer_dims(name)
        out = torch.empty((B, h, w, c), dtype=torch.float32, device=self.device)
        N.check(self.lib.b2t_extract(self.h, name.encode(), B, out.data_ptr(), _stream()))
        return out

    def profile_forward(self, frames: torch.Tensor):
        ms = (C.c_float * 23)()
        by = (C.c_double * 23)()
        dt = N.FRAME_U8 if frames.dtype == torch.uint8 else N.FRAME_F32
        N.check(self.lib.b2t_profile_forward(self.h, frames.data_ptr(), dt, frames.shape[0], ms, by, _stream()))
        return list(ms), list(by)

    # ---------------------------------------------------------------- decode
    def decode(self, logits: torch.Tensor, obj_threshold: float = 0.5, nms_threshold: float = 0.45,
               anchors: Optional[Sequence[float]] = None, max_boxes: Optional[int] = None, tag: str = ""):
        """decode_netout on the device.  logits (B,G,G,A,5+C) fp32 -> (boxes (B,max,8), counts (B)).  The result
        tensors are engine-owned scratch reused by the next decode of the same shape and `tag`."""
        B, gh, gw, nb, d = logits.shape
        anc = self._anchors if anchors is None else (C.c_float * (2 * nb))(*anchors)
        mb = max_boxes or gh * gw * nb
        boxes = self._buf(("boxes", B, mb, tag), (B, mb, 8), torch.float32)
        counts = self._buf(("counts", B, tag), (B,), torch.int32)
        N.check(self.lib.b2t_decode_nms(self.h, logits.data_ptr(), B, gh, gw, nb, d - 5, obj_threshold, nms_threshold,
                                        anc, boxes.data_ptr(), counts.data_ptr(), mb, _stream()))
        return boxes, counts

    def region_detect(self, logits: torch.Tensor, thresh: float, nms: float, orig_w: int, orig_h: int,
                      anchors: Optional[Sequence[float]] = None, max_dets: Optional[int] = None):
        """darknet region layer + get_network_boxes + do_nms_obj on the device."""
        B, gh, gw, nb, d = logits.shape
        anc = self._anchors if anchors is None else (C.c_float * (2 * nb))(*anchors)
        md = max_dets or gh * gw * nb
        dets = self._buf(("dets", B, md), (B, md, 8), torch.float32)
        counts = self._buf(("dcounts", B), (B,), torch.int32)
        N.check(self.lib.b2t_region_detect(self.h, logits.data_ptr(), B, gh, gw, nb, d - 5, thresh, nms, anc,
                                           orig_w, orig_h, self.image_size, self.image_size,
                                           dets.data_ptr(), counts.data_ptr(), md, _stream()))
        return dets, counts

    # ---------------------------------------------------------------- frame ingest
    def resize_frames(self, frames: torch.Tensor, size: Optional[int] = None) -> torch.Tensor:
        """cv2.resize(frame, (size, size)) on the device (KerasYOLO.py:526), bit-identical to OpenCV's INTER_LINEAR.
        frames: (B,H,W,3) uint8 CUDA tensor -> (B,size,size,3) uint8."""
        size = size or self.image_size
        if frames.dim() != 4 or frames.shape[3] != 3 or frames.dtype != torch.uint8 or not frames.is_cuda:
            raise ValueError("resize_frames expects a (B,H,W,3) uint8 CUDA tensor")
        frames = frames.contiguous()
        B, H, W = frames.shape[0], frames.shape[1], frames.shape[2]
        out = torch.empty((B, size, size, 3), dtype=torch.uint8, device=frames.device)
        N.check(self.lib.b2t_resize_frames(self.h, frames.data_ptr(), H, W, B, out.data_ptr(), size, size, _stream()))
        return out

    def letterbox_frames(self, frames: torch.Tensor, bgr: bool = False) -> torch.Tensor:
        """darknet ingest (load_image_color + letterbox_image, image.c:960-979,1442-1482) on the device:
        (B,H,W,3) uint8 frames of any size -> (B,S,S,3) float32 RGB in [0,1], what network_predict_image feeds."""
        if frames.dim() != 4 or frames.shape[3] != 3 or frames.dtype != torch.uint8 or not frames.is_cuda:
            raise ValueError("letterbox_frames expects a (B,H,W,3) uint8 CUDA tensor")
        frames = frames.contiguous()
        B, H, W = frames.shape[0], frames.shape[1], frames.shape[2]
        out = torch.empty((B, self.image_size, self.image_size, 3), dtype=torch.float32, device=frames.device)
        N.check(self.lib.b2t_letterbox_frames(self.h, frames.data_ptr(), H, W, B, 1 if bgr else 0, out.data_ptr(), _stream()))
        return out

    # ---------------------------------------------------------------- tracker helpers
    def pool_features(self, name: str, B: int, pool: str = "Global", chw_view: bool = False) -> torch.Tensor:
        h, w, c = self.layer_dims(name)
        n = c if pool == "Global" else (h // 4) * (w // 4) * c
        out = self._buf(("fv", name, B, pool), (B, n), torch.float32)
        N.check(self.lib.b2t_pool_features(self.h, name.encode(), B, 0 if pool == "Global" else 1,
                                           1 if chw_view else 0, out.data_ptr(), _stream()))
        return out

    def select_detection(self, dets: torch.Tensor, counts: torch.Tensor, frame_w: int, frame_h: int,
                         class_mask: Optional[torch.Tensor] = None, heat_size: int = 0):
        B, md, _ = dets.shape
        det_in = self._buf(("det_in", B), (B, 4), torch.float32)
        chosen = self._buf(("chosen", B), (B,), torch.int32)
        heat = self._buf(("heat", B, heat_size), (B, heat_size * heat_size), torch.float32) if heat_size else None
        N.check(self.lib.b2t_select_detection(self.h, dets.data_ptr(), counts.data_ptr(), md, B,
                                              class_mask.data_ptr() if class_mask is not None else None,
                                              frame_w, frame_h, det_in.data_ptr(), heat_size,
                                              heat.data_ptr() if heat is not None else None, chosen.data_ptr(), _stream()))
        return det_in, heat, chosen

    def heatmap_from_box(self, xywh: torch.Tensor, size: int = 32) -> torch.Tensor:
        n = xywh.shape[0]
        heat = torch.empty((n, size * size), dtype=torch.float32, device=self.device)
        N.check(self.lib.b2t_heatmap_from_box(self.h, xywh.data_ptr(), n, size, heat.data_ptr(), _stream()))
        return heat

    def box_from_heatmap(self, heat: torch.Tensor, size: int = 32, thresh: float = 0.75) -> torch.Tensor:
        n = heat.shape[0]
        rect = torch.empty((n, 4), dtype=torch.int32, device=self.device)
        N.check(self.lib.b2t_box_from_heatmap(self.h, heat.data_ptr(), n, size, thresh, rect.data_ptr(), _stream()))
        return rect

    # ---------------------------------------------------------------- callers after the path
    def draw_boxes(self, frames: torch.Tensor, rows: torch.Tensor, counts: torch.Tensor, color=(0, 255, 0)) -> torch.Tensor:
        """utils.draw_boxes' rectangles (cv2.rectangle, thickness 3) on (B,H,W,3) uint8 device frames, in place, for the
        first counts[b] decode rows of each frame.  Pixel-identical to OpenCV; text labels are a host job."""
        B, H, W, _ = frames.shape
        if frames.dtype != torch.uint8 or not frames.is_contiguous():
            raise ValueError("draw_boxes expects contiguous uint8 frames")
        N.check(self.lib.b2t_draw_boxes(self.h, frames.data_ptr(), B, H, W, rows.data_ptr(), counts.data_ptr(), rows.shape[1],
                                        int(color[0]), int(color[1]), int(color[2]), _stream()))
        return frames

    def overlap_scores(self, y_true: torch.Tensor, y_pred: torch.Tensor):
        """utils.overlap_score for n pairs of (x1,y1,x2,y2) float64 rows + average_overlap_score -> (scores (n), mean)."""
        n = y_true.shape[0]
        scores = torch.empty(n, dtype=torch.float64, device=self.device)
        mean = torch.empty(1, dtype=torch.float64, device=self.device)
        N.check(self.lib.b2t_overlap_scores(self.h, y_true.contiguous().data_ptr(), y_pred.contiguous().data_ptr(), n,
                                            scores.data_ptr(), mean.data_ptr(), _stream()))
        return scores, mean

    # ---------------------------------------------------------------- ConvLSTM (MultiObjDetTracker)
    def convlstm_reset(self) -> None:
        N.check(self.lib.b2t_convlstm_reset(self.h, _stream()))

    def convlstm_reset_slots(self, slot0: int, n: int = 1) -> None:
        N.check(self.lib.b2t_convlstm_reset_slots(self.h, slot0, n, _stream()))

    def convlstm_window(self, B: int, hard_sigmoid: bool = True) -> torch.Tensor:
        """Frames [0,B) of the last forward = B consecutive steps of one stream (state slot 0, carried over)."""
        out = self._buf(("trk", B), (B, self.grid, self.grid, self.n_box, 5 + self.n_class), torch.float32)
        N.check(self.lib.b2t_convlstm_window(self.h, B, out.data_ptr(), 1 if hard_sigmoid else 0, _stream()))
        return out

    def convlstm_sequence(self, S: int, T: int, slot0: int = 0, reset: bool = True,
                          hard_sigmoid: bool = True) -> torch.Tensor:
        """Frames [0,S*T) of the last forward = S streams x T consecutive steps (frame s*T + t), stream s in state
        slot slot0 + s -> tracker logits (S*T,G,G,A,5+C).  One batched input conv, T recurrent convs over the S
        streams, one batched head."""
        out = self._buf(("trk", S * T), (S * T, self.grid, self.grid, self.n_box, 5 + self.n_class), torch.float32)
        N.check(self.lib.b2t_convlstm_sequence(self.h, S, T, slot0, 1 if reset else 0, out.data_ptr(),
                                               1 if hard_sigmoid else 0, _stream()))
        return out

    @property
    def launches(self) -> int:
        """Kernels of this library launched so far: eager launches counted by the C side + the kernel nodes of
        every CUDA-graph replay (the C counter ticks once at capture time, each replay relaunches them)."""
        return self.lib.b2t_launch_count(self.h) + self._graph_launches

    def add_graph_launches(self, n: int) -> None:
        self._graph_launches += n


class LstmHead:
    """LSTM(units) + Dense(n_out, sigmoid) with device-resident state for ``max_streams`` streams."""

    def __init__(self, engine: DetectorEngine, n_feat: int, n_det: int, units: int, n_out: int, max_streams: int = 8):
        self.engine, self.lib = engine, engine.lib
        self.n_feat, self.n_det, self.units, self.n_out, self.max_streams = n_feat, n_det, units, n_out, max_streams
        h = C.c_void_p()
        N.check(self.lib.b2t_lstm_create(engine.h, n_feat, n_det, units, n_out, max_streams, C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.b2t_lstm_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_weights(self, w: Dict[str, np.ndarray]) -> None:
        a = [np.ascontiguousarray(w[k], dtype=np.float32) for k in
             ("kernel", "recurrent_kernel", "bias", "dense_kernel", "dense_bias")]
        if a[0].shape != (self.n_feat + self.n_det, 4 * self.units):
            raise ValueError(f"kernel shape {a[0].shape}")
        N.check(self.lib.b2t_lstm_set_weights(self.h, *[_np_ptr(x) for x in a], _stream()))

    def reset(self, stream_index: int = -1) -> None:
        N.check(self.lib.b2t_lstm_reset(self.h, stream_index, _stream()))

    def sequence(self, fv: torch.Tensor, det: torch.Tensor, reset: bool = True, hard_sigmoid: bool = True,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """fv (S,T,n_feat), det (S,T,n_det) contiguous -> y (S,T,n_out): T recurrent steps of S streams."""
        S, T = fv.shape[0], fv.shape[1]
        if not (fv.is_contiguous() and det.is_contiguous()):
            raise ValueError("sequence inputs must be contiguous")
        y = out if out is not None else torch.empty((S, T, self.n_out), dtype=torch.float32, device=fv.device)
        N.check(self.lib.b2t_lstm_sequence(self.h, fv.data_ptr(), det.data_ptr(), S, T, y.data_ptr(),
                                           1 if reset else 0, 1 if hard_sigmoid else 0, _stream()))
        return y

    def step(self, fv: torch.Tensor, det: torch.Tensor, hard_sigmoid: bool = True,
             out: Optional[torch.Tensor] = None, slot0: int = 0) -> torch.Tensor:
        """fv (S,n_feat), det (S,n_det): rows may be strided views (e.g. x[:, t] of an (S,T,F) tensor).  The S rows
        step the state slots [slot0, slot0+S); the other streams keep their (h, c)."""
        S = fv.shape[0]
        if fv.stride(-1) != 1 or det.stride(-1) != 1:
            raise ValueError("feature rows must be contiguous")
        y = out if out is not None else torch.empty((S, self.n_out), dtype=torch.float32, device=fv.device)
        N.check(self.lib.b2t_lstm_step_slots(self.h, slot0, fv.data_ptr(), fv.stride(0) if S > 1 else 0, det.data_ptr(),
                                             det.stride(0) if S > 1 else 0, S, y.data_ptr(),
                                             y.stride(0) if S > 1 else 0, 1 if hard_sigmoid else 0, _stream()))
        return y
