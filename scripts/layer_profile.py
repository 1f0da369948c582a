"""Developer tool run under gpurun: per-conv-layer time of the shipped engine next to each layer's tensor floor
(3 fp16 MMAs per product at the measured sustained peak) and HBM floor.  NB=batch, NC=classes."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from object_tracking_b200 import weights as W
from object_tracking_b200.engine import DetectorEngine

C = int(os.environ.get("NC", "80")); B = int(os.environ.get("NB", "16")); S = int(os.environ.get("SIZE", "416"))
try:
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pk = {"hbm_gbs": 6539.9, "bf16_tflops_sustained": 1344.4}
w = W.synthetic_yolo_weights(C, seed=0)
frames = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (B, S, S, 3), dtype=np.uint8)).cuda()
e = DetectorEngine(n_class=C, max_batch=B, image_size=S) if S != 416 else DetectorEngine(n_class=C, max_batch=B)
e.set_weights(w); e.finalize()
for _ in range(3):
    ms, by = e.profile_forward(frames)
acc = np.zeros(23)
R = 10
for _ in range(R):
    ms, by = e.profile_forward(frames)
    acc += np.array(ms)
acc /= R
# layer table: (cin, cout, k, out_hw before pool)
chans = [(3, 32, 3), (32, 64, 3), (64, 128, 3), (128, 64, 1), (64, 128, 3), (128, 256, 3), (256, 128, 1), (128, 256, 3),
         (256, 512, 3), (512, 256, 1), (256, 512, 3), (512, 256, 1), (256, 512, 3), (512, 1024, 3), (1024, 512, 1),
         (512, 1024, 3), (1024, 512, 1), (512, 1024, 3), (1024, 1024, 3), (1024, 1024, 3), (512, 64, 1), (1280, 1024, 3),
         (1024, 5 * (5 + C), 1)]
g = S // 32
hw = [S, S // 2, S // 4, S // 4, S // 4, S // 8, S // 8, S // 8] + [S // 16] * 5 + [g] * 7 + [2 * g, g, g]
tot_fl = 0.0
print(f"B={B} C={C} size={S}  peak {pk['bf16_tflops_sustained']} TF/s sustained")
print(f"{'layer':8s} {'ms':>8s} {'GFLOP':>8s} {'floor3x_ms':>10s} {'frac':>6s} {'MB':>8s}")
for i, ((ci, co, k), h) in enumerate(zip(chans, hw)):
    fl = 2.0 * B * h * h * ci * co * k * k
    tot_fl += fl
    floor = 3 * fl / (pk["bf16_tflops_sustained"] * 1e12) * 1e3
    print(f"conv_{i+1:<3d} {acc[i]:8.4f} {fl/1e9:8.2f} {floor:10.4f} {floor/acc[i] if acc[i] > 0 else 0:6.2f} {by[i]/1e6:8.1f}")
print(f"total ms {acc.sum():.4f}  per frame {acc.sum()/B*1e3:.1f} us; floor3x {3*tot_fl/(pk['bf16_tflops_sustained']*1e12)*1e3:.4f} ms")
