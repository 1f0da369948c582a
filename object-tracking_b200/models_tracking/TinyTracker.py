"""Drop-in for ``models_tracking/TinyTracker.py``: [pool(fv), detection bbox(4)] -> LSTM(lstm_units,
implementation=2) -> Dense(4, sigmoid) (TinyTracker.py:25-41), output = [cx, cy, w/2, h/2] image-relative
(utility/preprocessing.py:429-432)."""
from __future__ import annotations

from typing import Optional

import torch

from ..engine import LstmHead
from ..weights import synthetic_lstm_weights
from .BaseTracker import BaseTracker


class TinyTracker(BaseTracker):
    def __init__(self, config=None, tracker_weights: Optional[dict] = None, **kw):
        super(TinyTracker, self).__init__(config, **kw)
        self.LSTM_UNITS = self.config["model_tracker"]["lstm_units"]
        self.SEQUENCE_LENGTH = self.config["model_tracker"]["sequence_length"]
        self.n_det, self.n_out = 4, 4
        self._tracker_weights = tracker_weights
        self.load_tracker_model()

    def load_tracker_model(self):
        n_feat = self._n_feat()
        self.head = LstmHead(self.model_detector.engine, n_feat, self.n_det, self.LSTM_UNITS, self.n_out,
                             max_streams=self.max_streams)
        w = self._tracker_weights or synthetic_lstm_weights(n_feat + self.n_det, self.LSTM_UNITS, self.n_out, seed=1)
        self.head.set_weights(w)
        self.model_tracker = self.head

    def _tracker_inputs_from_state(self, B: int, W: int, H: int):
        fv, det_in, _, _ = self._decode_and_pool(B, W, H)
        return fv, det_in
