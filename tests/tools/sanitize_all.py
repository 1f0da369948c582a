"""One small-shape call of every kernel of libb200track.so, to be run under compute-sanitizer:

  compute-sanitizer --tool memcheck  python tests/tools/sanitize_all.py
  compute-sanitizer --tool racecheck python tests/tools/sanitize_all.py
  compute-sanitizer --tool synccheck python tests/tools/sanitize_all.py

Prints the list of kernels it exercised; the sanitizer's own summary follows on stderr."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from object_tracking_b200 import weights as W
from object_tracking_b200.engine import DetectorEngine, LstmHead

rng = np.random.default_rng(0)
C, U = 2, 64
w = W.synthetic_yolo_weights(C, seed=0)
done = []

def frames(b, s=416):
    return torch.from_numpy(rng.integers(0, 256, (b, s, s, 3), dtype=np.uint8)).cuda()

# conv_chain_kernel (batch 1), conv_pm / conv_halo_persist / conv_halo<big,small> / splitk_epilogue (batch 2, 3)
e = DetectorEngine(n_class=C, max_batch=3, convlstm_units=U, keep_prepool=True)
e.set_weights(w)
e.set_convlstm_weights(W.synthetic_multiobj_weights(C, U, seed=2))
e.finalize()
for b in (1, 2, 3):
    lg = e.forward(frames(b))
done += ["frames_to_c8", "conv_pm(mode 0,1)", "conv_chain", "conv_halo_persist", "conv_halo<big>", "conv_halo<small>", "splitk_epilogue"]
big = DetectorEngine(n_class=C, max_batch=20)                               # 20 frames: conv_19/20/22 run unsplit with two / three
big.set_weights(w); big.finalize()                                          # in-CTA accumulation passes (TMEM-parked sums)
big.forward(frames(20))
done.append("conv_halo<big> multi-pass")
del big
e.forward(torch.rand((1, 416, 416, 3), device="cuda"))                      # conv1_direct_kernel (float frames)
done.append("conv1_direct")
e.extract("norm_5", 3); e.extract("concat", 3)
done.append("planes_to_f32")
e.decode(e.logits(3), 0.5, 0.45)
e.region_detect(e.logits(3), 0.3, 0.45, 640, 480)
done += ["decode_nms<keras>", "decode_nms<darknet>"]
e.pool_features("conv_feat", 3, "Global"); e.pool_features("conv_feat", 3, "Max"); e.pool_features("conv_feat", 3, "Global", True)
done += ["pool_global", "pool_features"]
e.convlstm_sequence(1, 3, 0, True); e.convlstm_sequence(3, 1, 0, False); e.convlstm_window(2)
done += ["convlstm_gates", "conv_halo<big> (recurrent, accumulate)"]
dets, counts = e.region_detect(e.logits(3), 0.3, 0.45, 416, 416)
mask = torch.ones(C, dtype=torch.uint8, device="cuda")
e.select_detection(dets, counts, 416, 416, mask, 32)
e.heatmap_from_box(torch.rand((5, 4), device="cuda"), 32)
e.box_from_heatmap(torch.rand((5, 1024), device="cuda"), 32, 0.75)
done += ["select_detection", "heatmap_from_box", "box_from_heatmap"]
e.draw_boxes(frames(2), *e.decode(e.logits(2), 0.5, 0.45))
e.overlap_scores(torch.rand((9, 4), dtype=torch.float64, device="cuda"), torch.rand((9, 4), dtype=torch.float64, device="cuda"))
clip = torch.from_numpy(rng.integers(0, 256, (2, 3, 416, 416, 3), dtype=np.uint8)).cuda()
e.ingest_windows(clip[:, 1:2]); e.forward_ingested(2)
done += ["draw_boxes", "overlap_scores", "frames_to_c8 (strided ingest)"]
e.resize_frames(torch.from_numpy(rng.integers(0, 256, (2, 300, 400, 3), dtype=np.uint8)).cuda(), 416)
e.letterbox_frames(torch.from_numpy(rng.integers(0, 256, (2, 300, 400, 3), dtype=np.uint8)).cuda(), bgr=True)
done += ["resize_bilinear_u8", "letterbox_u8"]
head = LstmHead(e, 1024, 4, 512, 4, max_streams=3)
head.set_weights(W.synthetic_lstm_weights(1028, 512, 4, seed=1))
fv, det = torch.rand((3, 4, 1024), device="cuda"), torch.rand((3, 4, 4), device="cuda")
head.sequence(fv, det, reset=True)                                          # lstm_proj + lstm_seq (cooperative) + dense
head.step(fv[:, 0], det[:, 0]); head.step(fv[1:2, 1], det[1:2, 1], slot0=1)
done += ["lstm_proj", "lstm_seq", "lstm_gates", "dense_sigmoid"]
torch.cuda.synchronize()

# compat layer kernels (darknet ABI): letterbox, chw<->hwc, region activation
import ctypes as Ct, tempfile
from oracle import darknet_ref
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_darknet_abi as T
d = tempfile.mkdtemp()
cfg, wts = os.path.join(d, "y.cfg"), os.path.join(d, "y.weights")
darknet_ref.write_yolov2_cfg(cfg, 80, 416)
W.write_darknet_weights(wts, W.synthetic_yolo_weights(80, seed=0), 80)
lib = T.bind(Ct.CDLL(os.path.join(ROOT, "object-tracking_b200", "libb200track.so")))
net = lib.load_network(cfg.encode(), wts.encode(), 0)
im = lib.load_image_color(os.path.join(ROOT, "tests", "golden", "jpeg", "frame_500x300.jpg").encode(), 0, 0)
num = Ct.c_int(0)
lib.network_predict_image(net, im)
dd = lib.get_network_boxes(net, im.w, im.h, 0.25, 0.5, None, 0, Ct.byref(num))
lib.do_nms_obj(dd, num.value, 80, 0.45)
lib.network_extract_feat(net, 25)
lib.free_detections(dd, num.value)
lib.free_image(im)
done += ["compat: letterbox, chw_to_hwc, hwc_to_chw, region_activate", "reorg_gather (darknet semantics)"]
# the tiny graph (cfg/yolov2-tiny-voc.cfg): stride-1 max-pool kernel, 16-channel conv_1
tcfg, twts = os.path.join(d, "tiny.cfg"), os.path.join(d, "tiny.weights")
darknet_ref.write_tiny_weights(twts, 20, 1024, seed=3)
te = DetectorEngine(n_class=20, max_batch=2, semantics="darknet", graph="tiny", tiny_filters=1024)
te.load_darknet_weights(twts); te.finalize()
te.forward(frames(2)); te.forward(frames(1))
done += ["pool_s1 (tiny graph)"]
torch.cuda.synchronize()
print("exercised:", ", ".join(done))
