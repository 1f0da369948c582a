"""Developer check run under gpurun: per-layer error of both conv engines against the fp64 oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from object_tracking_b200 import weights as W
from object_tracking_b200.engine import DetectorEngine
from oracle import yolo_oracle

C = int(os.environ.get("NC", "2")); B = int(os.environ.get("NB", "2"))
w = W.synthetic_yolo_weights(C, seed=0)
rng = np.random.default_rng(1234)
frames = rng.integers(0, 256, (B, 416, 416, 3), dtype=np.uint8)
names = [f"norm_{i}" for i in range(1, 21)] + ["norm_22"]
t0 = time.time()
o = yolo_oracle.yolo_forward(yolo_oracle.normalize(frames), w, C, dtype=np.float64, want=names + ["concat"])
print(f"oracle fp64 {time.time()-t0:.1f}s", flush=True)
fr = torch.from_numpy(frames).cuda()
res = {}
for eng in sys.argv[1:] or ["simt", "tcgen05"]:
    e = DetectorEngine(n_class=C, max_batch=B, engine=eng, keep_prepool=True)
    e.set_weights(w); e.finalize()
    try:
        lg = e.forward(fr); torch.cuda.synchronize()
    except Exception as ex:
        print(eng, "FAILED", ex); continue
    print(f"== engine {eng}")
    for n in names + ["concat"]:
        key = n if n != "norm_22" else "feat"
        ref = o[key] if key in o else o[n]
        got = e.extract(n, B).cpu().numpy()
        err = np.abs(got - ref).max(); sc = np.abs(ref).max()
        print(f"  {n:8s} max|err| {err:.3e}  max|ref| {sc:.3f}  rel {err/sc:.2e}  nan={np.isnan(got).any()}")
    got = lg.cpu().numpy()
    err = np.abs(got - o["logits"]).max()
    print(f"  logits   max|err| {err:.3e} max|ref| {np.abs(o['logits']).max():.3f}")
    res[eng] = got
    ms, by = e.profile_forward(fr)
    ms, by = e.profile_forward(fr)
    print("  per-conv ms:", " ".join(f"{m:.3f}" for m in ms), " total", sum(ms))
if len(res) == 2:
    a, b = res.values()
    print("simt vs tcgen05 logits max diff", np.abs(a - b).max())
