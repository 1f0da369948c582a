"""Drop-in for ``models_tracking/TinyHeatmapTracker.py``: [pool(fv), 32x32 detection heat-map] ->
LSTM(lstm_units) -> Dense(heatmap_size**2, sigmoid) (TinyHeatmapTracker.py:26-48); heat-map in/out helpers
are utils.py:53-79 on the device."""
from __future__ import annotations

from typing import Optional

import torch

from ..engine import LstmHead
from ..weights import synthetic_lstm_weights
from .BaseTracker import BaseTracker


class TinyHeatmapTracker(BaseTracker):
    def __init__(self, config=None, tracker_weights: Optional[dict] = None, **kw):
        super(TinyHeatmapTracker, self).__init__(config, **kw)
        self.LSTM_UNITS = self.config["model_tracker"]["lstm_units"]
        self.SEQUENCE_LENGTH = self.config["model_tracker"]["sequence_length"]
        self.HEATMAP_SIZE = self.config["model_tracker"]["heatmap_size"]
        self.n_det = self.n_out = self.HEATMAP_SIZE * self.HEATMAP_SIZE
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
        fv, _, heat, _ = self._decode_and_pool(B, W, H, self.HEATMAP_SIZE)
        return fv, heat

    def rectangles(self, heat_out: torch.Tensor, thresh: float = 0.75) -> torch.Tensor:
        """(…, size*size) predicted heat-maps -> (…, 4) int32 [x1,y1,x2,y2] cells (utils.py:61-79)."""
        flat = heat_out.reshape(-1, self.n_out).contiguous()
        r = self.model_detector.engine.box_from_heatmap(flat, self.HEATMAP_SIZE, thresh)
        return r.view(*heat_out.shape[:-1], 4)
