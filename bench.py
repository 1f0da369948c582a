#!/usr/bin/env python
"""bench.py -- frames/sec of the 416x416 detect+track hot path (BASELINE.json metric) on N B200s.

Workload at N=1 = BASELINE.json configs[1]: TinyTracker = YOLOv2-416 (80 COCO classes, darknet semantics,
as models_detection/YOLO.py drives it) + LSTM(512) head, one stream, a synthetic 300-frame clip.  One
"step" = one batch of --windows (default 9) windows of model_tracker.sequence_length (4) consecutive frames of
the clip = 36 frames (windows are independent: the Keras LSTM is stateless across them, TinyTracker.py:26-37;
the detector is stateless per frame).  9 windows are chosen for the machine: 36 images x 8 output-channel tiles
= 288 work items on the 13x13 layers ~ 2 x 148 SMs; --windows 4 is the reference's config.json train.batch_size.
Per step: one batched detector pass, region decode + NMS, detection choice, feature pooling, the LSTM input
projection for all frames, 4 sequential recurrent steps over the windows, one batched Dense head.
--windows 1 gives the single-window latency case.  Steps are pipelined (TinyTracker.track_windows(pipeline=True): the
tracker tail of step i runs on a second stream under conv_1..8 of step i+1; --no-pipeline serialises them).
stdout carries exactly one JSON line (library chatter is redirected to stderr).  N>1: every rank runs its own
stream(s) -- independent units, no data-path collective; one NCCL broadcast of the packed weights at init.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--windows S] [--impl reference]

Prints ONE JSON line (see DESIGN.md section 6 for every field).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "frames_per_sec_416x416_detect_track"
UNIT = "frames/s"
N_CLASS, IMAGE, SEQ, CLIP = 80, 416, 4, 300


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe).  nvidia-smi takes a
    few hundred ms to emit its first line, so it is started before the warm-up and its timestamped samples are
    filtered to the timed window [mark_start, mark_end]."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line))

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        inside = [ln for (t, ln) in self.lines if self.t0 is not None and self.t0 <= t <= self.t1 + 0.03]
        window = "timed region"
        if len(inside) < 2:                                   # very short timed region: use every sample under load
            inside, window = [ln for (_, ln) in self.lines], "warm-up + timed region"
        sm, mx, reasons = [], [], set()
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------------ reference arm
def reference_runner():
    """The reference's own CPU implementation of the path: its darknet C library (oracle/_ref/libdarknet.so,
    compiled from /root/reference/darknet/src by oracle/Makefile) driven through the ctypes call sequence of
    models_detection/YOLO.py:140-170, + the numpy LSTM step (Keras is not installable here).  Falls back to
    the oracle port (torch-CPU forward) when the .so did not travel."""
    # all host threads, also under torchrun (it exports OMP_NUM_THREADS=1 for every rank)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(os.cpu_count())
    except OSError:
        pass
    from object_tracking_b200 import weights as W
    from oracle import darknet_ref, tracker_oracle, yolo_oracle
    w = W.synthetic_detector_weights(N_CLASS, seed=0)
    wl = {k: v.astype(np.float32) for k, v in W.synthetic_lstm_weights(1024 + 4, 512, 4, seed=1).items()}
    state = {"h": np.zeros((1, 512), np.float32), "c": np.zeros((1, 512), np.float32), "t": 0}

    def lstm(fv, det):
        if state["t"] % SEQ == 0:
            state["h"][:] = 0; state["c"][:] = 0
        state["t"] += 1
        y, state["h"], state["c"] = tracker_oracle.tracker_step(fv[None], det[None], state["h"], state["c"], wl)
        return y

    if darknet_ref.available():
        tmp = tempfile.mkdtemp()
        cfg, wpath = os.path.join(tmp, "yolov2.cfg"), os.path.join(tmp, "synthetic.weights")
        darknet_ref.write_yolov2_cfg(cfg, N_CLASS, IMAGE)
        W.write_darknet_weights(wpath, w, N_CLASS)
        net = darknet_ref.DarknetRef(cfg, wpath)

        def run(frame_u8):
            chw = np.ascontiguousarray(np.transpose(frame_u8.astype(np.float32) / np.float32(255.), (2, 0, 1)))
            net.predict(chw)
            boxes, obj, prob = net.detect(IMAGE, IMAGE, 0.5, 0.5, 0.45, N_CLASS)
            feat = net.extract(25).reshape(1024, -1).max(1)
            live = np.nonzero(prob.max(1) > 0)[0]
            det = np.zeros(4, np.float32)
            if live.size:
                j = live[np.argmax(prob[live].max(1))]
                det = (boxes[j] / np.float32(IMAGE)).astype(np.float32)
            return lstm(feat.astype(np.float32), det)
        return run, "reference", "libdarknet.so (reference C sources, -Ofast -fopenmp) + numpy LSTM step"

    from oracle import darknet_oracle

    def run(frame_u8):
        x = (frame_u8[None].astype(np.float32) / np.float32(255.))
        o = yolo_oracle.yolo_forward(x, w, N_CLASS, dtype=np.float32, mode="darknet", want=["norm_20"])
        logits = np.transpose(o["logits"].reshape(13, 13, -1), (2, 0, 1))
        region = darknet_oracle.region_forward(logits, N_CLASS)
        boxes, obj, prob = darknet_oracle.detect(region, IMAGE, IMAGE, IMAGE, IMAGE, 0.5, 0.45, N_CLASS)
        live = np.nonzero(prob.max(1) > 0)[0]
        det = np.zeros(4, np.float32)
        if live.size:
            j = live[np.argmax(prob[live].max(1))]
            det = (boxes[j] / np.float32(IMAGE)).astype(np.float32)
        return lstm(o["norm_20"][0].max(axis=(0, 1)).astype(np.float32), det)
    return run, "port", "oracle port: torch-CPU fp32 forward + numpy region decode/NMS + numpy LSTM step"


def workload_config(S: int, world: int) -> dict:
    """`config` of the JSON line -- the same object for the B200 arm and the reference arm."""
    return {"workload": "TinyTracker YOLOv2-416 C=80 (darknet semantics) + LSTM(512), 1 stream/window "
                        "per GPU x %d, synthetic 300-frame clip" % S,
            "frames_per_step": S * SEQ, "window": SEQ, "windows_per_step": S,
            "l2": "inputs cycle through a 156 MB clip per stream and the 204 MB weight blob is streamed "
                  "every step (both > 126 MB L2); no explicit flush",
            "weights": "random-init (reference ships none), seed 0", "parallelism": f"streams x{world}"}


def time_reference(steps: int, warmup: int):
    run, kind, what = reference_runner()
    rng = np.random.default_rng(1234)
    frames = rng.integers(0, 256, (max(1, min(8, steps + warmup)), IMAGE, IMAGE, 3), dtype=np.uint8)
    for i in range(warmup):
        run(frames[i % len(frames)])
    t0 = time.perf_counter()
    for i in range(steps):
        run(frames[(warmup + i) % len(frames)])
    dt = time.perf_counter() - t0
    return steps / dt, dt, kind, what


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    fps, dt, kind, what = time_reference(args.steps, args.warmup)
    cores = os.cpu_count()
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.windows, max(1, args.gpus)),
                           reference_step="1 frame per step: a bounded sample of the %d-frame step, same frames/s metric"
                                          % (args.windows * SEQ)),
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{args.steps} single frames after {args.warmup} warm-up; {what}"},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=JSON_OUT, flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ B200 arm
def main_b200(args):
    import torch
    import torch.distributed as dist
    from object_tracking_b200 import weights as W
    from object_tracking_b200.models_tracking.TinyTracker import TinyTracker

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    S, T = args.windows, SEQ
    cfg = {"model_detector": {"name": "YOLO", "config_file": "cfg/yolov2.cfg", "meta_file": "cfg/coco.data",
                              "weights_file": "yolov2.weights", "fv_layer": 25, "nms": 0.45, "thresh": 0.5, "hier_thresh": 0.5},
           "model_tracker": {"name": "TinyTracker", "lstm_units": 512, "sequence_length": T, "heatmap_size": 32},
           "train": {"cpu_only": 0, "dgpu_id": local, "tgpu_id": local, "pool": "Global", "batch_size": 4, "max_epochs": 0,
                     "tensorboard_dir": "logs/", "saved_model_dir": "models/", "classes": ["Person", "Car"]}}
    trk = TinyTracker(cfg, max_streams=S, detector_kwargs={"broadcast": world > 1, "rank": rank})
    eng = trk.model_detector.engine

    # synthetic clip(s): S streams x 300 frames, seeded per global stream id; resident copy + pinned host copy
    n_win = CLIP // T
    from object_tracking_b200.sharding import shard_streams
    gids = list(shard_streams(world * S, rank, world))         # global stream ids owned by this rank
    host = torch.empty((len(gids), CLIP, IMAGE, IMAGE, 3), dtype=torch.uint8, pin_memory=True)   # (S, 300, H, W, 3)
    for s_local, gid in enumerate(gids):                       # filled stream by stream: no second host copy
        rng = np.random.default_rng(1234 + gid)
        host[s_local].numpy()[...] = rng.integers(0, 256, (CLIP, IMAGE, IMAGE, 3), dtype=np.uint8)
    dev = host.cuda(non_blocking=True)
    torch.cuda.synchronize()

    def window(t, i):          # (S, T, H, W, 3) view of window i
        return t[:, (i % n_win) * T:(i % n_win) * T + T]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    PIPE = not args.no_pipeline          # tracker tail of step i overlaps conv_1..8 of step i+1 (BaseTracker.track_windows)
    for i in range(args.warmup):
        trk.track_windows(window(dev, i), pipeline=PIPE)
    eng.forward_events = []
    barrier()
    sampler.mark_start()
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = trk.track_windows(window(dev, args.warmup + i), pipeline=PIPE)
    if PIPE:
        torch.cuda.current_stream().wait_event(trk.tail_done)       # the last step's tail is inside the timed region
    e1.record()
    barrier()
    sampler.mark_end()
    launches = eng.launches - l0
    ms = e0.elapsed_time(e1)
    fwd_ms = sum(a.elapsed_time(b) for a, b in eng.forward_events) / max(1, len(eng.forward_events))
    eng.forward_events = None
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: through the plugin call (TinyTracker.track_windows) with HOST buffers: every step's frames travel
    # pinned host -> device inside the timed region and every step's result is read back to the host.  The copy of
    # step i+1 is issued on a second stream while step i computes (double-buffered device staging); the host reads
    # step i-1's boxes (async D2H into pinned memory + event) while step i runs, and the last step's before the clock stops.
    copy_stream = torch.cuda.Stream()
    stage = [torch.empty((S, T, IMAGE, IMAGE, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def upload(i, slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])                  # the step that last used this slot is done
            w = window(host, i)
            for s in range(S):                                      # S contiguous 2 MB pinned segments -> async DMA
                stage[slot][s].copy_(w[s], non_blocking=True)
            ready[slot].record(copy_stream)

    y_host = [torch.empty((S, T, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    y_done = [torch.cuda.Event(), torch.cuda.Event()]
    for slot in range(2):
        consumed[slot].record()
    for i in range(2):                                              # warm the path (graphs exist already)
        upload(i, i & 1)
        torch.cuda.current_stream().wait_event(ready[i & 1])
        y = trk.track_windows(stage[i & 1], pipeline=PIPE)
        consumed[i & 1].record()
        with torch.cuda.stream(trk.tail_stream if PIPE else torch.cuda.current_stream()):
            y_host[i & 1].copy_(y)
    barrier()
    t0 = time.perf_counter()
    upload(args.warmup, 0)
    checksum = 0.0
    for i in range(args.steps):
        slot = i & 1
        if i + 1 < args.steps:
            upload(args.warmup + i + 1, slot ^ 1)
        torch.cuda.current_stream().wait_event(ready[slot])
        y = trk.track_windows(stage[slot], pipeline=PIPE)
        consumed[slot].record()
        with torch.cuda.stream(trk.tail_stream if PIPE else torch.cuda.current_stream()):
            y_host[slot].copy_(y, non_blocking=True)                # D2H of this step's boxes into pinned memory
            y_done[slot].record()
        if i:                                                       # the host reads step i-1's boxes while step i runs
            y_done[slot ^ 1].synchronize()
            checksum += float(y_host[slot ^ 1][0, 0, 0])
    y_done[(args.steps - 1) & 1].synchronize()
    checksum += float(y_host[(args.steps - 1) & 1][0, 0, 0])
    barrier()
    e2e_s = time.perf_counter() - t0

    t_ms = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t_ms[0]), float(t_ms[1])
    frames_per_step = S * T
    total_frames = world * args.steps * frames_per_step

    if rank == 0:
        hbm, tflops, src = measured_peaks()
        B = frames_per_step
        bytes_per_fwd = W.forward_bytes(N_CLASS, IMAGE, batch=B) * B       # algorithmic, fp32-equivalent 4 B/elem
        flops_per_fwd = W.traffic_model(N_CLASS, IMAGE)["flops"] * B
        ach = bytes_per_fwd / (fwd_ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": total_frames / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f16x2-split operands, f32 accumulate", "data": "synthetic",
                "config": dict(workload_config(S, world),
                               pipeline="tracker tail of step i on a second stream under conv_1..8 of step i+1"
                               if PIPE else "serial"),
                "e2e": {"value": total_frames / (e2e_ms * 1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": frames_per_step * IMAGE * IMAGE * 3,
                        "d2h_bytes_per_step": frames_per_step * 4 * 4},
                "gpu_launches": int(launches),
                "clocks": clocks,
                # The conv stack is tensor-bound, not HBM-bound: parity needs 3 fp16 MMAs per product (DESIGN.md
                # section 4), so its tensor floor (3*flops / peak) is above its HBM floor at every batch size.
                # `achieved` = ALGORITHMIC flops (1x) / measured time; the HBM view is kept beside it.
                "roofline": {"bound": "tensor", "kernel": "YOLOv2 conv stack = frames_to_c8 + 3x conv_pm_kernel + 20x conv_halo_*"
                                                          "kernel (+ split-K epilogues), one CUDA-graph launch per step",
                             "achieved": flops_per_fwd / (fwd_ms * 1e-3) / 1e12, "peak": tflops, "unit": "TFLOP/s",
                             "frac": flops_per_fwd / (fwd_ms * 1e-3) / 1e12 / tflops,
                             # dram__bytes_read+write summed over the stack's 23 conv kernels of one 36-frame launch,
                             # ncu --set full capture in profiles/ (split-K epilogues and the u8->fp16 copy not included)
                             "traffic": 2.140e9 if B == 36 else None,
                             "peak_source": src + " (cuBLAS bf16, sustained)", "launch_ms": fwd_ms,
                             "algorithmic_flops_per_launch": flops_per_fwd,
                             "issued_tflops": 3 * flops_per_fwd / (fwd_ms * 1e-3) / 1e12,
                             # parity (bbox within 1e-3) needs 3 fp16 MMAs per product: the algorithmic fraction cannot
                             # exceed 1/3 of the tensor peak; frac_of_ceiling = issued MMA rate / peak
                             "ceiling_frac": 1.0 / 3.0,
                             "frac_of_ceiling": 3 * flops_per_fwd / (fwd_ms * 1e-3) / 1e12 / tflops,
                             "hbm": {"achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                                     "algorithmic_bytes_per_launch": bytes_per_fwd}}}
        if world == 1 and not args.no_cpu_baseline:
            fps, dt, kind, what = time_reference(40, 2)
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": os.cpu_count(), "kind": kind,
                                    "sample": f"40 frames (10 windows) after 2 warm-up frames, {dt:.1f} s; {what}"}
        print(json.dumps(line), file=JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--windows", type=int, default=9,
                    help="independent 4-frame windows per GPU per step (9 x 4 = 36 frames: 36 images x 8 cout tiles = 288 "
                         "work items ~ 2 x 148 SMs for the 13x13 layers)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="run the tracker tail of a step before the next step starts")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    # stdout carries exactly ONE JSON line: whatever libraries print there (NCCL's version banner, darknet's layer
    # table) is sent to stderr instead -- file descriptor 1 is pointed at stderr and the line is written to a duplicate
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.exit(main_reference(a) if a.impl == "reference" else main_b200(a))
