"""Developer tool (gpurun): conv-stack time per frame for small batches, one-kernel-per-layer schedule vs
conv_chain_kernel, CUDA graph replay, darknet semantics C=80.   python scripts/batch_sweep.py [size]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from object_tracking_b200 import weights as W
from object_tracking_b200.engine import DetectorEngine
S = int(sys.argv[1]) if len(sys.argv) > 1 else 416
w = W.synthetic_yolo_weights(80, seed=0)
for B in (1, 2, 4, 8, 16):
    row = []
    for chain in (-1, 64):
        e = DetectorEngine(n_class=80, max_batch=B, image_size=S, semantics="darknet", chain_max_batch=chain)
        e.set_weights(w); e.finalize()
        fr = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (B, S, S, 3), dtype=np.uint8)).cuda()
        for _ in range(3):
            e.forward(fr)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            e.forward(fr)
        for _ in range(5):
            g.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            g.replay()
        b.record(); torch.cuda.synchronize()
        row.append(a.elapsed_time(b) / 50 * 1e3 / B)
        del e, g
    print(f"size {S} B={B:2d}: per-layer {row[0]:7.1f} us/frame   chain {row[1]:7.1f} us/frame", flush=True)
