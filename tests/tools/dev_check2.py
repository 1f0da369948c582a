"""gpurun helper: per-layer error of the engines vs the fp64 oracle + per-conv timings at a given batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from object_tracking_b200 import weights as W
from object_tracking_b200.engine import DetectorEngine
from oracle import yolo_oracle
C = int(os.environ.get("NC", "2")); B = int(os.environ.get("NB", "2")); TB = int(os.environ.get("TB", "4"))
w = W.synthetic_yolo_weights(C, seed=0)
frames = np.random.default_rng(1234).integers(0, 256, (max(B, TB), 416, 416, 3), dtype=np.uint8)
names = [f"norm_{i}" for i in range(1, 21)] + ["concat"]
o = yolo_oracle.yolo_forward(yolo_oracle.normalize(frames[:B]), w, C, dtype=np.float64, want=names)
fr = torch.from_numpy(frames).cuda()
for eng in sys.argv[1:] or ["tcgen05", "tcgen05_tile"]:
    e = DetectorEngine(n_class=C, max_batch=max(B, TB), engine=eng, keep_prepool=os.environ.get("KEEP", "1") == "1")
    e.set_weights(w); e.finalize()
    lg = e.forward(fr[:B]).cpu().numpy(); torch.cuda.synchronize()
    print(f"== engine {eng}")
    bad = []
    for n in (names if os.environ.get("KEEP", "1") == "1" else []):
        got = e.extract(n, B).cpu().numpy(); ref = o[n]
        rel = np.abs(got - ref).max() / np.abs(ref).max()
        if rel > 2e-5 or np.isnan(got).any(): bad.append((n, float(rel)))
    feat = e.extract("conv_feat", B).cpu().numpy()
    print("  layers over 2e-5:", bad, " feat rel", np.abs(feat - o["feat"]).max() / np.abs(o["feat"]).max(),
          " logits err", np.abs(lg - o["logits"]).max())
    e.forward(fr[:TB]); e.forward(fr[:TB])
    ms, by = e.profile_forward(fr[:TB]); ms, by = e.profile_forward(fr[:TB])
    print(f"  B={TB} per-conv us:", " ".join(f"{1e3*m:.0f}" for m in ms), " total ms", round(sum(ms), 3))
