"""examples/c_host.c: a plain C99 program on the C-ABI (no Python, no torch in the process).  CPU suite: both headers are
valid C and the example compiles and links against libb200track.so.  GPU suite: its detections, logits and pooled feature
equal what the Python plugin path produces on the same weights file and frames (same kernels: bit for bit)."""
import os
import re
import subprocess

import numpy as np
import pytest

from object_tracking_b200 import _native as N, weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build(out):
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CUDA, "include"),
           os.path.join(ROOT, "examples", "c_host.c"), "-o", out, "-L" + os.path.dirname(N.LIB_PATH), "-lb200track",
           "-L" + os.path.join(CUDA, "lib64"), "-lcudart", "-Wl,-rpath," + os.path.dirname(N.LIB_PATH)]
    return subprocess.run(cmd, capture_output=True, text=True)


def test_headers_are_plain_c_and_the_example_links(tmp_path):
    for h in ("b200track.h", "darknet_compat.h"):
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", h)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    r = _build(str(tmp_path / "c_host"))
    assert r.returncode == 0, r.stderr
    assert os.path.exists(tmp_path / "c_host")


@pytest.mark.gpu
def test_c_host_equals_the_python_path(tmp_path):
    import torch
    from object_tracking_b200.engine import DetectorEngine
    B, C = 3, 80
    wts, raw, exe = str(tmp_path / "planted.weights"), str(tmp_path / "frames.u8"), str(tmp_path / "c_host")
    W.write_darknet_weights(wts, W.synthetic_detector_weights(C, seed=0), C)
    frames = np.random.default_rng(11).integers(0, 256, (B, 416, 416, 3), dtype=np.uint8)
    frames.tofile(raw)
    assert _build(exe).returncode == 0
    r = subprocess.run([exe, wts, raw, str(B), str(C)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("frame")]
    assert len(lines) == B
    e = DetectorEngine(n_class=C, max_batch=B, semantics="darknet")
    e.load_darknet_weights(wts)
    e.finalize()
    lg = e.forward(torch.from_numpy(frames).cuda())
    dets, counts = e.region_detect(lg, 0.5, 0.45, 416, 416)
    fv = e.pool_features("norm_20", B, "Global").cpu().numpy()
    dets, counts, lg = dets.cpu().numpy(), counts.cpu().numpy(), lg.cpu().numpy()
    n_with = 0
    for i, line in enumerate(lines):
        m = re.match(r"frame (\d+): (-?\d+) detections; first: (.*); logit (\S+) fv (\S+)", line)
        assert m, line
        assert int(m.group(1)) == i and int(m.group(2)) == int(counts[i]), (line, counts[i])
        first = [float(v) for v in m.group(3).split()]
        if counts[i] > 0:
            n_with += 1
            assert np.array_equal(np.float32(first[:6]), dets[i, 0, :6]) and int(first[6]) == int(dets[i, 0, 6]), (line, dets[i, 0])
        assert np.float32(float(m.group(4))) == lg[i].reshape(-1)[0]
        assert np.float32(float(m.group(5))) == fv[i, 0]
    assert n_with >= 1                                              # the planted detector fires on these frames
