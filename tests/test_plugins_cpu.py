"""CPU suite, part 3: host-side logic of the plugin layer that needs no GPU -- reference-name aliasing, config
defaults, class constants, stream sharding, and the world_size-2 gloo path of the weight broadcast."""
import importlib
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_module_names_resolve():
    import object_tracking_b200 as b2t
    b2t.install_reference_aliases()
    for name in ("TinyTracker", "TinyHeatmapTracker", "MultiObjDetTracker"):
        cls = getattr(importlib.import_module("models_tracking." + name), name)     # trainer.py:12-14
        assert cls.__name__ == name
    from models_detection.KerasYOLO import KerasYOLO
    from utility.utils import BoundBox, bbox_iou, decode_netout, normalize  # noqa: F401
    assert KerasYOLO.OBJ_THRESHOLD == 0.5 and KerasYOLO.NMS_THRESHOLD == 0.45 and KerasYOLO.BOX == 5
    assert len(KerasYOLO.LABELS) == 80 and KerasYOLO.MAX_BOX_PER_IMAGE == 50
    assert KerasYOLO.ANCHORS[:2] == [0.57273, 0.677385] and KerasYOLO.weight_path == 'darknet/yolov2.weights'
    from models_tracking.MultiObjDetTracker import MultiObjDetTracker
    assert MultiObjDetTracker.CLASS == 12 and MultiObjDetTracker.SEQUENCE_LENGTH == 4


def test_host_helpers_follow_reference_semantics():
    from object_tracking_b200.utility.utils import BoundBox, bbox_iou, softmax, WeightReader
    a = BoundBox(0.5, 0.5, 0.2, 0.2, 0.9, np.array([0.1, 0.8]))
    b = BoundBox(0.55, 0.5, 0.2, 0.2, 0.9, np.array([0.7, 0.2]))
    assert a.get_label() == 1 and abs(a.get_score() - 0.8) < 1e-12
    assert abs(bbox_iou(a, b) - (0.15 * 0.2) / (0.08 - 0.15 * 0.2)) < 1e-9
    x = np.array([[0.0, -150.0], [1.0, 2.0]])
    s = softmax(x)                                   # global max / global-min rescale of utils.py:262-270
    assert np.allclose(s.sum(-1), 1.0) and s[0, 1] > np.exp(-150.0)
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".weights") as f:
        np.arange(10, dtype=np.float32).tofile(f.name)
        r = WeightReader(f.name)
        assert list(r.read_bytes(3)) == [4.0, 5.0, 6.0]          # 4-word header skipped (utils.py:138-148)


def test_config_defaults_match_reference_config_json():
    from object_tracking_b200.models_detection._common import load_config
    c = load_config(None) if not os.path.exists("config.json") else load_config({})
    c = load_config(None)
    assert c["model_detector"]["fv_layer"] == 25 and c["model_detector"]["nms"] == 0.45
    assert c["model_tracker"]["sequence_length"] == 4 and c["model_tracker"]["lstm_units"] == 512
    assert c["train"]["pool"] == "Global" and c["train"]["classes"] == ["Person", "Car"]


def test_stream_sharding_is_a_partition():
    from object_tracking_b200.sharding import shard_streams
    for n_streams, world in ((32, 8), (64, 8), (5, 2), (1, 4), (7, 3)):
        seen = []
        for r in range(world):
            mine = shard_streams(n_streams, r, world)
            assert mine == sorted(mine)
            seen += mine
        assert sorted(seen) == list(range(n_streams))
        sizes = [len(shard_streams(n_streams, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    assert shard_streams(32, 3, 8) == [3, 11, 19, 27]             # stream i -> GPU i mod n


@pytest.mark.timeout(180)
def test_weight_broadcast_protocol_world_size_2_gloo(tmp_path):
    """The N>1 init path on CPU/gloo: rank 0 packs the blob, the other rank receives it with one broadcast and
    ends up bit-identical; streams are partitioned without overlap."""
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, hashlib
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch, torch.distributed as dist
        from object_tracking_b200.sharding import shard_streams, broadcast_blob
        from object_tracking_b200 import weights as W
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        n = 1 << 20
        if rank == 0:
            w = W.synthetic_lstm_weights(64, 32, 4, seed=5)
            blob = torch.from_numpy(np.concatenate([v.ravel() for v in w.values()]).view(np.uint8).copy())
            blob = torch.cat([blob, torch.zeros(n - blob.numel(), dtype=torch.uint8)])
        else:
            blob = torch.empty(n, dtype=torch.uint8)
        broadcast_blob(blob, src=0)
        digest = hashlib.sha1(blob.numpy().tobytes()).hexdigest()
        mine = shard_streams(5, rank, world)
        out = [None] * world
        dist.all_gather_object(out, (digest, mine))
        if rank == 0:
            assert out[0][0] == out[1][0], "blob differs after broadcast"
            assert sorted(out[0][1] + out[1][1]) == [0, 1, 2, 3, 4]
            print("OK", out)
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
                       capture_output=True, text=True, timeout=170, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "OK" in r.stdout


def test_overlap_metrics_match_reference_source():
    """utility/utils.py:82-110 (tracking metric helpers, callers after the path): same numbers as the reference's own
    functions exec'd from its source when /root/reference is present, fixed known answers otherwise."""
    import importlib
    U = importlib.import_module("object_tracking_b200.utility.utils")
    a, b = (10., 20., 50., 80.), (30., 40., 70., 100.)
    assert abs(U.overlap_score(a, b) - (20. * 40.) / (40. * 60. * 2 - 20. * 40.)) < 1e-12
    assert U.average_overlap_score([a, a], [a, b]) == (1.0 + U.overlap_score(a, b)) / 2
    ref_path = "/root/reference/utility/utils.py"
    if os.path.exists(ref_path):
        src = open(ref_path).read().split("\n")
        ns = {}
        exec(compile("\n".join(src[81:110]), "ref_utils_metrics", "exec"), ns)     # overlap_score, average_overlap_score
        rng = np.random.default_rng(0)
        for _ in range(50):
            t = np.sort(rng.uniform(0, 100, 4)); p = np.sort(rng.uniform(0, 100, 4))
            t = (t[0], t[1], t[2], t[3]); p = (p[0], p[1], p[2], p[3])
            assert U.overlap_score(t, p) == ns["overlap_score"](t, p)


def test_item_decode_magic_multiplier_is_exact():
    """conv_pm_kernel decodes its work items with host-made magic multipliers (api.cu magic_for, conv_pm.cu
    decode_item): q = umulhi(n, ceil(2^32 / d)) must equal n // d for every item index the kernels can see
    (n < 2^32 / d).  The arithmetic is restated here and checked exhaustively on the divisors in use and on
    random ones."""
    rng = np.random.default_rng(0)
    divisors = [2, 3, 4, 5, 7, 13, 14, 52, 104, 152, 208, 304, 1456] + [int(x) for x in rng.integers(2, 5000, 40)]
    for d in divisors:
        m = ((1 << 32) + d - 1) // d
        assert m < (1 << 32)
        hi = min((1 << 32) // d, 1 << 22)
        n = np.concatenate([np.arange(0, min(hi, 200000), dtype=np.uint64),
                            rng.integers(0, hi, 100000).astype(np.uint64),
                            np.array([hi - 1], dtype=np.uint64)])
        q = (n * np.uint64(m)) >> np.uint64(32)
        assert np.array_equal(q, n // np.uint64(d)), d


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the reference's own CPU path, the arm the driver runs beside ours): exactly one JSON line
    on stdout with the contract's keys, the same metric / unit / config as the B200 arm, and its cpu_baseline / e2e blocks."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]
