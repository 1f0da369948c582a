"""CPU oracle for the detect-and-track hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; nothing under ``object-tracking_b200/``
does.  It restates, on the CPU, what the reference computes on the path (file:line cited per
function) so the CUDA results can be checked against it.

Parity pinning: the reference ships NO golden vectors, tests or weights for this path
(SURVEY.md section 4 / 8c), and its Keras/TF half cannot run here (python 3.12, no
tensorflow).  What *is* pinned, by ``oracle/make_golden.py`` run in the build container:

* ``decode_oracle.decode_netout``  == the reference's own ``utility/utils.py:208-270`` exec'd
  from ``/root/reference`` on thousands of seeded/adversarial tensors (bit-exact);
* ``yolo_oracle.yolo_forward(mode="darknet")`` == the reference's darknet C library
  (``oracle/_ref/libdarknet.so`` built from ``/root/reference/darknet/src``) on the same
  ``.weights`` file and input (conv/BN/leaky/maxpool/reorg/route/region);
* ``darknet_oracle`` box decode + ``do_nms_obj`` == the same library's outputs.

* ``tracker_oracle.generate_heatmap_feat / generate_rectangle_from_heatmap`` == the reference's own
  ``utility/utils.py:53-79`` exec'd from ``/root/reference`` (3 000 cases; 96 outputs committed);
* ``ingest_oracle`` == the installed OpenCV's ``cv2.resize`` (bit-exact);
* the JPEG decoder's fixtures == ``load_image_color`` of the reference C library (bit-exact).

The Keras-only semantics (BN eps 1e-3, tf.space_to_depth ordering, LSTM / ConvLSTM2D gate
equations) are restated from the published Keras 2 definitions: for those rows parity is
"unpinned" in the sense of the task statement and DESIGN.md says so.  Their structure (gate order,
weight layouts, state update) is cross-checked against independent implementations --
``torch.nn.LSTMCell`` and a torch ``conv2d`` composition -- in ``tests/test_oracle_cpu.py``.
"""
