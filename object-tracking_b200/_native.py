"""ctypes binding of libb200track.so (include/b200track.h).  The library is the product: if it is not
built, or the machine has no sm_100 GPU, every compute call raises -- there is no Python/CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2T_USE_DEV_LIB=1 loads the developer build (make DEV=1: cross-check engines, B2T_* overrides) when it exists
LIB_PATH = os.path.join(_HERE, "libb200track.so")
if os.environ.get("B2T_USE_DEV_LIB") == "1" and os.path.exists(os.path.join(_HERE, "libb200track_dev.so")):
    LIB_PATH = os.path.join(_HERE, "libb200track_dev.so")

SEM_KERAS, SEM_DARKNET = 0, 1
ENGINE_TCGEN05, ENGINE_SIMT, ENGINE_TCGEN05_TILE = 0, 1, 2
FRAME_U8, FRAME_F32 = 0, 1


class B2TError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("image_h", C.c_int), ("image_w", C.c_int), ("n_class", C.c_int), ("max_batch", C.c_int),
                ("semantics", C.c_int), ("bn_eps", C.c_float), ("engine", C.c_int), ("device", C.c_int),
                ("convlstm_units", C.c_int), ("reserved", C.c_int * 7)]


_vp, _fp, _ip = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)

# name -> (restype, argtypes); every symbol include/b200track.h declares
SIGNATURES = {
    "b2t_last_error": (C.c_char_p, []),
    "b2t_version": (C.c_int, []),
    "b2t_dev_build": (C.c_int, []),
    "b2t_create": (C.c_int, [C.POINTER(Config), C.POINTER(_vp)]),
    "b2t_destroy": (None, [_vp]),
    "b2t_weight_bytes": (C.c_size_t, [_vp]),
    "b2t_workspace_bytes": (C.c_size_t, [_vp]),
    "b2t_bind_memory": (C.c_int, [_vp, _vp, _vp]),
    "b2t_set_conv_weights": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b2t_load_darknet_weights": (C.c_int, [_vp, C.c_char_p]),
    "b2t_set_convlstm_weights": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "b2t_finalize": (C.c_int, [_vp, C.c_int, _vp]),
    "b2t_yolo_forward": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "b2t_yolo_forward_range": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "b2t_ingest_frames": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_longlong, _vp]),
    "b2t_logits": (_vp, [_vp]),
    "b2t_extract": (C.c_long, [_vp, C.c_char_p, C.c_int, _vp, _vp]),
    "b2t_layer_dims": (C.c_int, [_vp, C.c_char_p, _ip, _ip, _ip]),
    "b2t_decode_nms": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                 _fp, _vp, _vp, C.c_int, _vp]),
    "b2t_region_detect": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                    _fp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp]),
    "b2t_lstm_create": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "b2t_lstm_destroy": (None, [_vp]),
    "b2t_lstm_set_weights": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b2t_lstm_reset": (C.c_int, [_vp, C.c_int, _vp]),
    "b2t_lstm_step": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, C.c_int, _vp, C.c_int, C.c_int, _vp]),
    "b2t_lstm_step_slots": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_int, _vp, C.c_int, C.c_int, _vp]),
    "b2t_lstm_sequence": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, C.c_int, C.c_int, _vp]),
    "b2t_pool_features": (C.c_int, [_vp, C.c_char_p, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "b2t_resize_frames": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int, _vp]),
    "b2t_letterbox_frames": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "b2t_heatmap_from_box": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "b2t_select_detection": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp,
                                       _vp, _vp]),
    "b2t_box_from_heatmap": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_float, _vp, _vp]),
    "b2t_convlstm_reset": (C.c_int, [_vp, _vp]),
    "b2t_convlstm_window": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp]),
    "b2t_convlstm_reset_slots": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "b2t_convlstm_sequence": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp]),
    "b2t_draw_boxes": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "b2t_overlap_scores": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, _vp, _vp]),
    "b2t_graph_begin": (C.c_int, [_vp, _vp]),
    "b2t_graph_end": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "b2t_graph_launch": (C.c_int, [_vp, _vp]),
    "b2t_graph_destroy": (None, [_vp]),
    "b2t_broadcast_weights": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "b2t_launch_count": (C.c_long, [_vp]),
    "b2t_profile_forward": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _fp, C.POINTER(C.c_double), _vp]),
}

_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise B2TError(f"{LIB_PATH} is not built: run `python __graft_entry__.py` (build) first; "
                           "there is no fallback path")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _LIB = handle
    return _LIB


def check(rc: int) -> int:
    if rc < 0:
        raise B2TError(lib().b2t_last_error().decode("utf-8", "replace") or f"libb200track error {rc}")
    return rc
