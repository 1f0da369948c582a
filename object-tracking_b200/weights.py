"""Layer table of the YOLOv2 detector, darknet ``.weights`` IO and seeded synthetic weights.

Host-side only (numpy).  What it mirrors in the reference:

* layer table            -- models_detection/KerasYOLO.py:277-400 (conv_1..conv_23, the five
                            MaxPooling2D, the conv_21 skip + space_to_depth + concatenate)
* ``read_darknet_weights`` -- utility/utils.py:138-148 (WeightReader) +
                            models_detection/KerasYOLO.py:244-274 (init_weights) and the file
                            header rules of darknet/src/parser.c:1214-1226
* ``write_darknet_weights`` -- inverse of the above (darknet/src/parser.c:1149-1198 order:
                            biases, [scales, rolling_mean, rolling_variance], weights)

No weights ship with the reference, so benches and tests use
``synthetic_yolo_weights`` -- random-init weights of the same architecture.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

ANCHORS = [0.57273, 0.677385, 1.87446, 2.06253, 3.33843, 5.47434, 7.88282, 3.52778, 9.77052, 9.16828]
N_BOX = 5
BN_EPS_KERAS = 1e-3          # keras BatchNormalization default epsilon


@dataclass(frozen=True)
class ConvSpec:
    """One conv_k of KerasYOLO.load_model (index = Keras name conv_<index>)."""
    index: int
    ksize: int
    cin: int
    cout: int
    bn: bool            # BatchNormalization + LeakyReLU(0.1) follow; else bias + linear
    pool: bool          # MaxPooling2D(2,2) follows
    src: str            # "prev" | "skip" (conv_13 pre-pool output) | "concat" ([s2d(conv_21), conv_20])


def yolo_layer_table(n_class: int = 80) -> List[ConvSpec]:
    """conv_1..conv_23 in weight-file order (KerasYOLO.py:277-400)."""
    out_ch = N_BOX * (4 + 1 + n_class)
    t = [
        (3, 3, 32, True), (3, 32, 64, True), (3, 64, 128, False), (1, 128, 64, False),
        (3, 64, 128, True), (3, 128, 256, False), (1, 256, 128, False), (3, 128, 256, True),
        (3, 256, 512, False), (1, 512, 256, False), (3, 256, 512, False), (1, 512, 256, False),
        (3, 256, 512, True), (3, 512, 1024, False), (1, 1024, 512, False), (3, 512, 1024, False),
        (1, 1024, 512, False), (3, 512, 1024, False), (3, 1024, 1024, False), (3, 1024, 1024, False),
    ]
    specs = [ConvSpec(i + 1, k, ci, co, True, p, "prev") for i, (k, ci, co, p) in enumerate(t)]
    specs.append(ConvSpec(21, 1, 512, 64, True, False, "skip"))
    specs.append(ConvSpec(22, 3, 1280, 1024, True, False, "concat"))
    specs.append(ConvSpec(23, 1, 1024, out_ch, False, False, "prev"))
    return specs


def n_params(n_class: int = 80) -> int:
    n = 0
    for s in yolo_layer_table(n_class):
        n += s.ksize * s.ksize * s.cin * s.cout + (4 * s.cout if s.bn else s.cout)
    return n


def traffic_model(n_class: int = 80, image_size: int = 416) -> Dict[str, float]:
    """Algorithmic element counts of one forward pass in fused-minimum form:
    W = conv weights + conv_23 bias, R = activation elements read (each conv input once, conv_1's input
    as stored), Wr = activation elements written (pooled where a pool follows; conv_13 both full and
    pooled; the concat written in place), flops = 2*MACs."""
    h = image_size
    w_el = r_el = wr_el = 0
    flops = 0.0
    for s in yolo_layer_table(n_class):
        if s.src == "skip":
            hh = image_size // 16
        elif s.src == "concat" or s.index == 23:
            hh = image_size // 32
        else:
            hh = h
        px = hh * hh
        w_el += s.ksize * s.ksize * s.cin * s.cout + (0 if s.bn else s.cout)
        r_el += px * s.cin
        out = px * s.cout
        wr_el += (out // 4 if s.pool else out) + (out if s.index == 13 else 0)
        flops += 2.0 * px * s.ksize * s.ksize * s.cin * s.cout
        if s.pool and s.src == "prev":
            h //= 2
    return {"W": w_el, "R": r_el, "Wr": wr_el, "flops": flops}


def forward_bytes(n_class: int = 80, image_size: int = 416, batch: int = 1, elem_bytes: int = 4) -> float:
    """Algorithmic HBM bytes PER FRAME at `batch` frames per weight pass: s*(W/B + R + Wr)."""
    t = traffic_model(n_class, image_size)
    return elem_bytes * (t["W"] / batch + t["R"] + t["Wr"])


# --------------------------------------------------------------------------------------
# synthetic weights
# --------------------------------------------------------------------------------------

def synthetic_yolo_weights(n_class: int = 80, seed: int = 0, head_obj_bias: float = -2.0,
                           head_gains=(0.13, 0.05, 0.40, 0.80),
                           class_bias: Optional[Dict[int, float]] = None) -> Dict[str, np.ndarray]:
    """Random-init weights of the KerasYOLO architecture, Keras layouts.

    kernel_k : (kh, kw, Cin, Cout) float32   gamma_k/beta_k/mean_k/var_k : (Cout,)   bias_23 : (Cout,)
    He-scaled kernels keep activations O(1..10) through 22 layers; the head (conv_23) columns are
    scaled per entry type (``head_gains`` = xy, wh, objectness, class; conv_feat has rms ~7 on
    random frames) and the objectness bias shifted so that t_xy ~ N(0,1), t_wh ~ N(0,0.4) and a few
    dozen anchors pass the 0.5 threshold on random frames.
    ``class_bias`` {class index: offset} is added to those classes' conv_23 bias for every anchor ("planted"
    classes: a random head emits the same few classes on every random frame, none of them the tracker's
    allowed ones -- see ``synthetic_detector_weights``).
    """
    rng = np.random.default_rng(seed)
    w: Dict[str, np.ndarray] = {}
    for s in yolo_layer_table(n_class):
        fan_in = s.ksize * s.ksize * s.cin
        # gain for LeakyReLU(0.1): sqrt(2/(1+0.01))
        std = np.sqrt(2.0 / 1.01 / fan_in) if s.bn else np.sqrt(1.0 / fan_in)
        ker = rng.standard_normal((s.ksize, s.ksize, s.cin, s.cout), dtype=np.float32) * np.float32(std)
        if not s.bn:
            d = 5 + n_class
            col = np.empty(d, np.float32)
            col[0:2], col[2:4], col[4], col[5:] = head_gains
            ker *= np.tile(col, N_BOX)[None, None, None, :]
        w[f"kernel_{s.index}"] = ker
        if s.bn:
            w[f"gamma_{s.index}"] = rng.uniform(0.5, 1.5, s.cout).astype(np.float32)
            w[f"beta_{s.index}"] = (0.1 * rng.standard_normal(s.cout)).astype(np.float32)
            w[f"mean_{s.index}"] = (0.1 * rng.standard_normal(s.cout)).astype(np.float32)
            w[f"var_{s.index}"] = rng.uniform(0.5, 1.5, s.cout).astype(np.float32)
        else:
            b = (0.1 * rng.standard_normal(s.cout)).astype(np.float32)
            d = 5 + n_class
            b.reshape(N_BOX, d)[:, 4] += np.float32(head_obj_bias)
            for c, off in (class_bias or {}).items():
                b.reshape(N_BOX, d)[:, 5 + int(c)] += np.float32(off)
            w[f"bias_{s.index}"] = b
    return w


PLANTED_CLASS_BIAS = {0: 10.0, 2: 9.0}     # COCO "person", "car" = config.json train.classes (config.json:39)


def synthetic_detector_weights(n_class: int = 80, seed: int = 0) -> Dict[str, np.ndarray]:
    """The random-init detector the YOLO plugin, bench.py and the tracker tests share when no
    ``darknet/yolov2.weights`` exists: ``synthetic_yolo_weights`` with the tracker's allowed classes planted, so
    that on random frames most (not all) frames carry person / car detections, several per frame -- the
    detection choice of utility/preprocessing.py:434-449 (class filter, top probability, zeros if none) is then
    exercised on both branches."""
    bias = {c: v for c, v in PLANTED_CLASS_BIAS.items() if c < n_class} if n_class >= 80 else None
    return synthetic_yolo_weights(n_class, seed=seed, class_bias=bias)


# --------------------------------------------------------------------------------------
# darknet .weights file
# --------------------------------------------------------------------------------------

def write_darknet_weights(path: str, w: Dict[str, np.ndarray], n_class: int = 80,
                          major: int = 0, minor: int = 1, revision: int = 0, seen: int = 0) -> None:
    """Emit a darknet weights file: header (3 x int32 + seen), then per conv
    biases(beta) [scales(gamma) rolling_mean rolling_variance] weights[Cout][Cin][kh][kw]."""
    with open(path, "wb") as f:
        f.write(struct.pack("<iii", major, minor, revision))
        if major * 10 + minor >= 2 and major < 1000 and minor < 1000:
            f.write(struct.pack("<Q", seen))
        else:
            f.write(struct.pack("<i", seen))
        for s in yolo_layer_table(n_class):
            if s.bn:
                for k in ("beta", "gamma", "mean", "var"):
                    f.write(np.ascontiguousarray(w[f"{k}_{s.index}"], dtype="<f4").tobytes())
            else:
                f.write(np.ascontiguousarray(w[f"bias_{s.index}"], dtype="<f4").tobytes())
            k_oihw = np.transpose(w[f"kernel_{s.index}"], (3, 2, 0, 1))
            f.write(np.ascontiguousarray(k_oihw, dtype="<f4").tobytes())


def read_darknet_weights(path: str, n_class: int = 80, strict: bool = True) -> Dict[str, np.ndarray]:
    """Parse a darknet weights file into Keras-layout arrays.

    Header handling follows darknet/src/parser.c:1214-1226 (``seen`` is int32 for v0.1 files and
    size_t from v0.2 on); the reference's own WeightReader (utils.py:138-148) always skips 16
    bytes, i.e. it is only right for v0.1 files -- both give the same result there.
    """
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size < 16:
        raise ValueError(f"{path}: too short to be a darknet weights file")
    major, minor, _rev = struct.unpack("<iii", raw[:12].tobytes())
    off = 12 + (8 if (major * 10 + minor >= 2 and major < 1000 and minor < 1000) else 4)
    body = raw[off:]
    if body.size % 4:
        raise ValueError(f"{path}: payload is not a whole number of float32")
    flt = body.view("<f4")
    pos = 0

    def take(n: int) -> np.ndarray:
        nonlocal pos
        if pos + n > flt.size:
            raise ValueError(f"{path}: truncated (need {pos + n} floats, file holds {flt.size})")
        a = flt[pos:pos + n]
        pos += n
        return np.array(a, dtype=np.float32)

    w: Dict[str, np.ndarray] = {}
    for s in yolo_layer_table(n_class):
        if s.bn:
            w[f"beta_{s.index}"] = take(s.cout)
            w[f"gamma_{s.index}"] = take(s.cout)
            w[f"mean_{s.index}"] = take(s.cout)
            w[f"var_{s.index}"] = take(s.cout)
        else:
            w[f"bias_{s.index}"] = take(s.cout)
        k = take(s.cout * s.cin * s.ksize * s.ksize).reshape(s.cout, s.cin, s.ksize, s.ksize)
        w[f"kernel_{s.index}"] = np.ascontiguousarray(np.transpose(k, (2, 3, 1, 0)))
    if strict and pos != flt.size:
        raise ValueError(f"{path}: {flt.size - pos} trailing floats (wrong class count?)")
    return w


# --------------------------------------------------------------------------------------
# tracker heads
# --------------------------------------------------------------------------------------

def synthetic_lstm_weights(n_in: int, units: int, n_out: int, seed: int = 1) -> Dict[str, np.ndarray]:
    """Keras-2 LSTM + Dense weights: kernel (n_in,4u), recurrent_kernel (u,4u), bias (4u,) in gate
    order i,f,c,o; dense_kernel (u,n_out), dense_bias (n_out,)  (TinyTracker.py:36-37)."""
    rng = np.random.default_rng(seed)
    g = lambda *s, sc: (rng.standard_normal(s) * sc).astype(np.float32)
    b = np.zeros(4 * units, np.float32)
    b[units:2 * units] = 1.0                      # keras unit_forget_bias
    b += g(4 * units, sc=0.05)
    return {
        "kernel": g(n_in, 4 * units, sc=1.0 / np.sqrt(n_in)),
        "recurrent_kernel": g(units, 4 * units, sc=1.0 / np.sqrt(units)),
        "bias": b,
        "dense_kernel": g(units, n_out, sc=2.0 / np.sqrt(units)),
        "dense_bias": g(n_out, sc=0.1),
    }


def synthetic_convlstm_weights(cin: int, units: int, n_out: int, seed: int = 2, n_class: int = 0,
                               head_class_gain: float = 1.0, head_obj_bias: float = 0.0,
                               head_wh_gain: float = 1.0) -> Dict[str, np.ndarray]:
    """Keras-2 ConvLSTM2D(units,3x3,same) + 1x1 head: kernel (3,3,cin,4u), recurrent_kernel
    (3,3,u,4u), bias (4u,), gate order i,f,c,o; head_kernel (1,1,u,n_out), head_bias
    (MultiObjDetTracker.py:176-183).  With ``n_class`` the head's class columns are scaled by ``head_class_gain``
    and its objectness bias shifted by ``head_obj_bias`` (n_out = 5*(5+n_class)): a random head never clears the
    0.5 class-score threshold, the shaped one yields MOT17-like box counts per frame;
    ``head_wh_gain`` scales the t_w / t_h columns (boxes of a random head are otherwise many images wide)."""
    rng = np.random.default_rng(seed)
    g = lambda *s, sc: (rng.standard_normal(s, dtype=np.float32) * np.float32(sc))
    b = np.zeros(4 * units, np.float32)
    b[units:2 * units] = 1.0
    b += g(4 * units, sc=0.05)
    w = {
        "kernel": g(3, 3, cin, 4 * units, sc=1.0 / np.sqrt(9 * cin)),
        "recurrent_kernel": g(3, 3, units, 4 * units, sc=1.0 / np.sqrt(9 * units)),
        "bias": b,
        "head_kernel": g(1, 1, units, n_out, sc=2.0 / np.sqrt(units)),
        "head_bias": g(n_out, sc=0.1),
    }
    if n_class:
        d = 5 + n_class
        assert n_out == N_BOX * d
        w["head_kernel"].reshape(units, N_BOX, d)[:, :, 5:] *= np.float32(head_class_gain)
        w["head_bias"].reshape(N_BOX, d)[:, 5:] *= np.float32(head_class_gain)
        w["head_bias"].reshape(N_BOX, d)[:, 4] += np.float32(head_obj_bias)
        w["head_kernel"].reshape(units, N_BOX, d)[:, :, 2:4] *= np.float32(head_wh_gain)
        w["head_bias"].reshape(N_BOX, d)[:, 2:4] *= np.float32(head_wh_gain)
    return w


def synthetic_multiobj_weights(n_class: int, units: int = 512, seed: int = 2) -> Dict[str, np.ndarray]:
    """The random-init tracker head MultiObjDetTracker, its tests and bench.py share when no trained checkpoint exists."""
    n_out = N_BOX * (5 + n_class)
    return synthetic_convlstm_weights(n_out + 1024, units, n_out, seed=seed, n_class=n_class, head_class_gain=4.0,
                                      head_obj_bias=-0.3, head_wh_gain=0.25)


# --------------------------------------------------------------------------------------
# tracker / detector checkpoints (BaseTracker.py:74-80 ModelCheckpoint files, MultiObjDetTracker.py:291-293)
# --------------------------------------------------------------------------------------

def _leaf(path: str) -> str:
    return path.rstrip("/").split("/")[-1].split(":")[0]


def load_checkpoint_arrays(path: str) -> Dict[str, np.ndarray]:
    """Every array of a checkpoint file keyed by its path: Keras ``.hdf5`` / ``.h5`` (read with hdf5_lite -- h5py is not
    needed) or ``.npz`` (numpy; what ``save_tracker_checkpoint`` writes, or a conversion of a Keras file made where
    h5py exists: ``np.savez(out, **{name: f[name][()] for name in dataset_names})``)."""
    if path.endswith(".npz"):
        with np.load(path) as z:
            return {k: np.asarray(z[k]) for k in z.files}
    from .hdf5_lite import read_hdf5
    return read_hdf5(path)


def lstm_weights_from_arrays(arrays: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Pick the LSTM + Dense weights of TinyTracker / TinyHeatmapTracker (TinyTracker.py:36-37: ``recurrent_layer`` +
    TimeDistributed(Dense ``output``)) out of a checkpoint by structure: the group that owns a 2-D ``recurrent_kernel``
    is the LSTM; the Dense kernel is the 2-D ``kernel`` whose rows equal the LSTM's units."""
    if {"kernel", "recurrent_kernel", "bias", "dense_kernel", "dense_bias"} <= set(arrays):
        return {k: np.asarray(arrays[k], np.float32) for k in ("kernel", "recurrent_kernel", "bias", "dense_kernel", "dense_bias")}
    rec = [k for k, v in arrays.items() if _leaf(k) == "recurrent_kernel" and v.ndim == 2]
    if len(rec) != 1:
        raise ValueError(f"expected exactly one 2-D recurrent_kernel in the checkpoint, found {len(rec)}")
    grp = rec[0].rsplit("/", 1)[0]
    units = arrays[rec[0]].shape[0]
    pick = lambda leaf: next(v for k, v in arrays.items() if k.startswith(grp + "/") and _leaf(k) == leaf)
    out = {"kernel": pick("kernel"), "recurrent_kernel": arrays[rec[0]], "bias": pick("bias")}
    dense = [k for k, v in arrays.items() if _leaf(k) == "kernel" and v.ndim == 2 and v.shape[0] == units and not k.startswith(grp + "/")]
    if len(dense) != 1:
        raise ValueError(f"expected exactly one Dense kernel with {units} rows, found {len(dense)}")
    dg = dense[0].rsplit("/", 1)[0]
    out["dense_kernel"] = arrays[dense[0]]
    out["dense_bias"] = next(v for k, v in arrays.items() if k.startswith(dg + "/") and _leaf(k) == "bias")
    if out["kernel"].shape[1] != 4 * units or out["bias"].shape != (4 * units,):
        raise ValueError("LSTM kernel / bias shapes do not match its recurrent kernel")
    return {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in out.items()}


def convlstm_weights_from_arrays(arrays: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """ConvLSTM2D ``tconv_lstm`` + 1x1 head ``tconv_2`` of MultiObjDetTracker (MultiObjDetTracker.py:176-183), by
    structure: the 4-D ``recurrent_kernel`` marks the ConvLSTM group; the head is the 1x1 4-D kernel fed by its units."""
    if {"kernel", "recurrent_kernel", "bias", "head_kernel", "head_bias"} <= set(arrays):
        return {k: np.asarray(arrays[k], np.float32) for k in ("kernel", "recurrent_kernel", "bias", "head_kernel", "head_bias")}
    rec = [k for k, v in arrays.items() if _leaf(k) == "recurrent_kernel" and v.ndim == 4]
    if len(rec) != 1:
        raise ValueError(f"expected exactly one 4-D recurrent_kernel in the checkpoint, found {len(rec)}")
    grp = rec[0].rsplit("/", 1)[0]
    units = arrays[rec[0]].shape[2]
    pick = lambda leaf: next(v for k, v in arrays.items() if k.startswith(grp + "/") and _leaf(k) == leaf)
    head = [k for k, v in arrays.items() if _leaf(k) == "kernel" and v.ndim == 4 and v.shape[:3] == (1, 1, units)
            and not k.startswith(grp + "/") and not _is_detector_conv(k)]
    if len(head) != 1:
        raise ValueError(f"expected exactly one 1x1 head kernel fed by {units} channels, found {len(head)}")
    hg = head[0].rsplit("/", 1)[0]
    out = {"kernel": pick("kernel"), "recurrent_kernel": arrays[rec[0]], "bias": pick("bias"), "head_kernel": arrays[head[0]],
           "head_bias": next(v for k, v in arrays.items() if k.startswith(hg + "/") and _leaf(k) == "bias")}
    return {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in out.items()}


def _is_detector_conv(path: str) -> bool:
    import re
    return re.search(r"(^|/)conv_\d+(/|$)", path) is not None


def detector_weights_from_arrays(arrays: Dict[str, np.ndarray], n_class: int) -> Optional[Dict[str, np.ndarray]]:
    """The YOLOv2 layers ``conv_k`` / ``norm_k`` (KerasYOLO.py:277-400 names) of a Keras checkpoint, or None when the
    file does not hold all of them (a tracker-only checkpoint)."""
    import re
    w: Dict[str, np.ndarray] = {}
    names = {"kernel": "kernel", "gamma": "gamma", "beta": "beta", "moving_mean": "mean", "moving_variance": "var", "bias": "bias"}
    for k, v in arrays.items():
        m = re.search(r"(?:^|/)(conv|norm)_(\d+)/[^/]*$", k) or re.search(r"(?:^|/)(conv|norm)_(\d+)/(?:[^/]+/)*[^/]+$", k)
        if not m or _leaf(k) not in names:
            continue
        w[f"{names[_leaf(k)]}_{int(m.group(2))}"] = np.ascontiguousarray(v, dtype=np.float32)
    for s in yolo_layer_table(n_class):
        need = [f"kernel_{s.index}"] + ([f"{n}_{s.index}" for n in ("gamma", "beta", "mean", "var")] if s.bn else [f"bias_{s.index}"])
        if any(n not in w for n in need):
            return None
        if w[f"kernel_{s.index}"].shape != (s.ksize, s.ksize, s.cin, s.cout):
            raise ValueError(f"conv_{s.index}: kernel shape {w[f'kernel_{s.index}'].shape} does not match {n_class} classes")
    return w


def save_tracker_checkpoint(path: str, w: Dict[str, np.ndarray]) -> None:
    """``.npz`` with this package's own keys (the dict the tracker classes take as ``tracker_weights``)."""
    np.savez(path, **{k: np.asarray(v, np.float32) for k, v in w.items()})


def latest_checkpoint(prefix: str) -> Optional[str]:
    """BaseTracker.py:74-80 writes ``<prefix>-CHKPNT-<epoch>-<val_loss>.hdf5``: the file of the highest epoch (an
    ``.npz`` conversion next to it is preferred), or None."""
    import glob
    import re
    best = None
    for f in glob.glob(prefix + "-CHKPNT-*"):
        m = re.search(r"-CHKPNT-(\d+)-", f)
        if not m or not f.endswith((".hdf5", ".h5", ".npz")):
            continue
        key = (int(m.group(1)), f.endswith(".npz"))
        if best is None or key > best[0]:
            best = (key, f)
    return best[1] if best else None
