"""CPU suite: on-disk formats next to the hot path: the minimal HDF5 reader for Keras
checkpoints against h5py-written files of the reference tree, checkpoint -> weight-dict mapping, .npz round trip."""
import os

import numpy as np
import pytest

from object_tracking_b200 import weights as W
from object_tracking_b200.hdf5_lite import HDF5Error, read_hdf5

REF_H5 = "/root/reference/py-faster-rcnn/caffe-fast-rcnn/src/caffe/test/test_data"


@pytest.mark.skipif(not os.path.exists(REF_H5), reason="/root/reference not present")
def test_hdf5_reader_on_h5py_written_files():
    """Contents known from the reference's generate_sample_data.py: contiguous float32 datasets and gzip-compressed
    chunked uint8 / float32 datasets, old-style groups (what h5py's default settings and Keras produce)."""
    total = 8 * 10 * 6 * 5
    data = np.arange(total).reshape(10, 8, 6, 5).astype("float32")
    label = (1 + np.arange(10)[:, None]).astype("float32")
    a = read_hdf5(os.path.join(REF_H5, "sample_data.h5"))
    assert sorted(a) == ["/data", "/label", "/label2"]
    assert np.array_equal(a["/data"], data) and np.array_equal(a["/label"], label) and np.array_equal(a["/label2"], label + 1)
    b = read_hdf5(os.path.join(REF_H5, "sample_data_2_gzip.h5"))
    assert b["/label"].dtype == np.uint8 and np.array_equal(b["/label"], label.astype("uint8"))
    assert np.array_equal(b["/data"], data + total) and np.array_equal(b["/label2"], (label + 1).astype("uint8"))
    c = read_hdf5(os.path.join(REF_H5, "solver_data.h5"))
    assert c["/data"].shape == (8, 3, 10, 10) and c["/targets"].shape == (8, 1) and abs(float(c["/data"].std()) - 1) < 0.2


def test_hdf5_reader_rejects_other_files(tmp_path):
    p = str(tmp_path / "x.hdf5")
    open(p, "wb").write(b"not hdf5 at all" * 10)
    with pytest.raises(HDF5Error):
        read_hdf5(p)


def test_checkpoint_mapping_and_npz_roundtrip(tmp_path):
    w = W.synthetic_lstm_weights(1028, 512, 4, seed=3)
    # Keras names as model.save() stores them (TinyTracker.py:36-37: LSTM 'recurrent_layer', TimeDistributed(Dense))
    keras = {"/model_weights/recurrent_layer/recurrent_layer/kernel:0": w["kernel"],
             "/model_weights/recurrent_layer/recurrent_layer/recurrent_kernel:0": w["recurrent_kernel"],
             "/model_weights/recurrent_layer/recurrent_layer/bias:0": w["bias"],
             "/model_weights/time_distributed_2/time_distributed_2/kernel:0": w["dense_kernel"],
             "/model_weights/time_distributed_2/time_distributed_2/bias:0": w["dense_bias"],
             "/optimizer_weights/Adam/iterations:0": np.array(7, np.int64)}
    got = W.lstm_weights_from_arrays(keras)
    assert all(np.array_equal(got[k], w[k]) for k in w)
    p = str(tmp_path / "TinyTracker-CHKPNT-03-0.41.npz")
    W.save_tracker_checkpoint(p, w)
    assert all(np.array_equal(W.lstm_weights_from_arrays(W.load_checkpoint_arrays(p))[k], w[k]) for k in w)
    open(str(tmp_path / "TinyTracker-CHKPNT-12-0.30.hdf5"), "wb").write(b"x")
    open(str(tmp_path / "TinyTracker-CHKPNT-02-0.90.hdf5"), "wb").write(b"x")
    assert W.latest_checkpoint(str(tmp_path / "TinyTracker")).endswith("CHKPNT-12-0.30.hdf5")
    assert W.latest_checkpoint(str(tmp_path / "Nothing")) is None
    # MultiObjDetTracker: ConvLSTM2D 'tconv_lstm' + TimeDistributed(Conv2D 'tconv_2') + the detector's conv_k / norm_k
    C, U = 2, 64
    wl = W.synthetic_multiobj_weights(C, U, seed=2)
    wd = W.synthetic_yolo_weights(C, seed=0)
    arrays = {"/model_weights/tconv_lstm/tconv_lstm/kernel:0": wl["kernel"],
              "/model_weights/tconv_lstm/tconv_lstm/recurrent_kernel:0": wl["recurrent_kernel"],
              "/model_weights/tconv_lstm/tconv_lstm/bias:0": wl["bias"],
              "/model_weights/timedist_tconv2/timedist_tconv2/kernel:0": wl["head_kernel"],
              "/model_weights/timedist_tconv2/timedist_tconv2/bias:0": wl["head_bias"]}
    for s in W.yolo_layer_table(C):
        base = f"/model_weights/timedist_bbox/conv_{s.index}"
        arrays[f"{base}/kernel:0"] = wd[f"kernel_{s.index}"]
        if s.bn:
            nb = f"/model_weights/timedist_bbox/norm_{s.index}"
            for keras_name, ours in (("gamma", "gamma"), ("beta", "beta"), ("moving_mean", "mean"), ("moving_variance", "var")):
                arrays[f"{nb}/{keras_name}:0"] = wd[f"{ours}_{s.index}"]
        else:
            arrays[f"{base}/bias:0"] = wd[f"bias_{s.index}"]
    got = W.convlstm_weights_from_arrays(arrays)
    assert all(np.array_equal(got[k], wl[k]) for k in wl)
    det = W.detector_weights_from_arrays(arrays, C)
    assert det is not None and all(np.array_equal(det[k], wd[k]) for k in wd)
    assert W.detector_weights_from_arrays({k: v for k, v in arrays.items() if "conv_7" not in k}, C) is None
