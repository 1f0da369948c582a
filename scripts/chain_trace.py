"""Developer tool (gpurun, dev library): per-layer clock stamps of conv_chain_kernel's CTA 0.
  B2T_USE_DEV_LIB=1 B2T_TRACE_CONV=100 NB=1 python scripts/chain_trace.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from object_tracking_b200 import weights as W
from object_tracking_b200.engine import DetectorEngine
B = int(os.environ.get("NB", "1")); C = int(os.environ.get("NC", "80")); S = int(os.environ.get("SIZE", "416"))
e = DetectorEngine(n_class=C, max_batch=B, image_size=S, semantics="darknet")
e.set_weights(W.synthetic_yolo_weights(C, seed=0)); e.finalize()
fr = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (B, S, S, 3), dtype=np.uint8)).cuda()
for _ in range(3):
    e.forward(fr)
torch.cuda.synchronize()
