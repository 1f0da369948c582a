"""CPU suite: the JPEG decoder behind the compat layer's load_image_color (csrc/jpeg_decode.cu; the reference decodes
through stb_image, darknet/src/image.c:1442-1482, and models_detection/YOLO.py:141 always passes a .jpg path) is
bit-exact with the reference library (sequential and progressive files): committed outputs of oracle/_ref/libdarknet.so on the fixtures of
tests/golden/jpeg/, the library itself side by side where it is present, and the reference's own data/*.jpg when
/root/reference exists.  load_image_color is host code: no GPU needed."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from oracle import darknet_ref
from object_tracking_b200 import _native as N

GOLD = os.path.join(os.path.dirname(__file__), "golden")


class IMAGE(C.Structure):          # YOLO.py:20-24
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("c", C.c_int), ("data", C.POINTER(C.c_float))]


def _bind(path):
    lib = C.CDLL(path)
    lib.load_image_color.restype = IMAGE
    lib.load_image_color.argtypes = [C.c_char_p, C.c_int, C.c_int]
    lib.free_image.argtypes = [IMAGE]
    return lib


def _load(lib, path, w=0, h=0):
    im = lib.load_image_color(path.encode(), w, h)
    if not im.data:
        return None
    a = np.ctypeslib.as_array(im.data, shape=(im.c, im.h, im.w)).copy()
    lib.free_image(im)
    return a


def test_jpeg_decoder_matches_reference_outputs():
    ours = _bind(N.LIB_PATH)
    z = np.load(os.path.join(GOLD, "jpeg_cases.npz"))
    assert len(z.files) >= 15 and any(n.startswith("prog") for n in z.files)
    for name in z.files:
        got = _load(ours, os.path.join(GOLD, "jpeg", name))
        assert got is not None, (name, N.lib().b2t_last_error())
        ref = (z[name].astype(np.float64) / 255.0).astype(np.float32)      # load_image_stb: (float)byte / 255.
        assert got.shape == ref.shape and got.dtype == np.float32, name
        assert np.array_equal(got, ref), (name, float(np.abs(got - ref).max() * 255))


@pytest.mark.skipif(not darknet_ref.available(), reason="oracle/_ref/libdarknet.so not present")
def test_jpeg_decoder_side_by_side_with_reference_library(tmp_path):
    import cv2
    ours, ref = _bind(N.LIB_PATH), _bind(darknet_ref.LIB_PATH)
    files = sorted(glob.glob(os.path.join(GOLD, "jpeg", "*.jpg"))) + sorted(glob.glob("/root/reference/darknet/data/*.jpg"))
    rng = np.random.default_rng(7)
    for i, q in enumerate((5, 30, 100)):                               # fresh files: extreme qualities, odd sizes
        for prog in (0, 1):
            p = str(tmp_path / f"r{i}_{prog}.jpg")
            cv2.imwrite(p, rng.integers(0, 256, (41 + 13 * i, 29 + 17 * i, 3), dtype=np.uint8),
                        [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_PROGRESSIVE, prog])
            files.append(p)
    for p in files:
        a, b = _load(ref, p), _load(ours, p)
        assert b is not None and a.shape == b.shape and np.array_equal(a, b), p
    # the (w, h) form resizes like load_image (image.c:1467-1471 -> resize_image)
    p = os.path.join(GOLD, "jpeg", "s420.jpg")
    a, b = _load(ref, p, 64, 48), _load(ours, p, 64, 48)
    assert a.shape == b.shape == (3, 48, 64) and np.abs(a - b).max() < 1e-6


def test_unsupported_and_corrupt_files_fail_without_exiting(tmp_path):
    import cv2
    ours = _bind(N.LIB_PATH)
    p = str(tmp_path / "x.png")
    cv2.imwrite(p, np.zeros((4, 4, 3), np.uint8))
    assert _load(ours, p) is None and b"decodes JPEG" in N.lib().b2t_last_error()
    good = open(os.path.join(GOLD, "jpeg", "s444.jpg"), "rb").read()
    p = str(tmp_path / "cut.jpg")
    open(p, "wb").write(good[:len(good) // 3])                          # truncated scan: decodes what is there, no crash
    _load(ours, p)
    p = str(tmp_path / "junk.jpg")
    open(p, "wb").write(b"\xff\xd8" + bytes(range(256)) * 4)
    assert _load(ours, p) is None
    # binary PPM still works
    p = str(tmp_path / "a.ppm")
    img = np.arange(5 * 7 * 3, dtype=np.uint8).reshape(5, 7, 3)
    open(p, "wb").write(b"P6\n7 5\n255\n" + img.tobytes())
    got = _load(ours, p)
    assert np.array_equal(got, (np.transpose(img, (2, 0, 1)).astype(np.float64) / 255.0).astype(np.float32))
