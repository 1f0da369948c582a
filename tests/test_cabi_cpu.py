"""CPU suite, part 2: the C-ABI shared library loads and exports every symbol include/b200track.h declares,
argument validation works without a GPU, and compute calls fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from object_tracking_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:b2t|dk)_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = N.lib()
    names = _declared("b200track.h")
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b200track.h but not exported"
        assert n in N.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.b2t_version() >= 100


def test_create_validates_and_sizes():
    lib = N.lib()
    cfg = N.Config()
    cfg.image_h = cfg.image_w = 416
    cfg.n_class, cfg.max_batch, cfg.bn_eps = 80, 1, 1e-3
    h = C.c_void_p()
    assert lib.b2t_create(C.byref(cfg), C.byref(h)) == 0
    # 50.9 M weights as fp16 hi/lo pairs = 4 bytes per weight (+ padding of conv_2's Cin to 64)
    assert 203e6 < lib.b2t_weight_bytes(h) < 206e6
    assert lib.b2t_workspace_bytes(h) > 30e6
    lib.b2t_destroy(h)
    cfg.image_h = 400
    assert lib.b2t_create(C.byref(cfg), C.byref(h)) < 0
    assert b"multiple of 32" in lib.b2t_last_error()


def test_compute_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = N.lib()
    cfg = N.Config()
    cfg.image_h = cfg.image_w = 416
    cfg.n_class, cfg.max_batch, cfg.bn_eps = 2, 1, 1e-3
    h = C.c_void_p()
    assert lib.b2t_create(C.byref(cfg), C.byref(h)) == 0
    assert lib.b2t_finalize(h, 0, None) < 0                       # no device -> error, not a fallback
    assert lib.b2t_yolo_forward(h, None, 0, 1, None, None) < 0
    assert b"not finalized" in lib.b2t_last_error()
    lib.b2t_destroy(h)
    from object_tracking_b200.engine import DetectorEngine
    with pytest.raises(N.B2TError):
        DetectorEngine(n_class=2)


def test_library_exports_the_reference_darknet_abi():
    """Every symbol models_detection/YOLO.py:58-119 binds (and include/darknet_compat.h declares)."""
    import ctypes
    lib = ctypes.CDLL(N.LIB_PATH)
    src = open(os.path.join(ROOT, "include", "darknet_compat.h")).read()
    bound_by_yolo_py = ["network_width", "network_height", "cuda_set_device", "get_network_boxes", "free_detections",
                        "free_ptrs", "load_network", "do_nms_obj", "free_image", "get_metadata", "load_image_color",
                        "rgbgr_image", "network_predict_image", "network_extract_feat", "layer_dims"]
    for n in bound_by_yolo_py + ["network_predict", "make_image", "free_network"]:
        assert re.search(r"\b%s\s*\(" % n, src), f"{n} not declared in darknet_compat.h"
        assert hasattr(lib, n), f"{n} not exported"
