"""Autoregressive rollout of the forecast step with the state kept on the device.

Mirrors the reference's flat rollout loop (applications/rollout_to_netcdf.py:269-310): ``y = model(x)``,
then ``x <- update_x(x, forcing, y)`` (datasets/gen_2/channel_utils.py:253-291): prognostic channels come
from the prediction, dynamic-forcing channels from the data stream, static channels are carried.
Only single-frame inputs (history_len == 1) are supported, like ``build_channel_layout`` (channel_utils.py:205-211).
"""

from __future__ import annotations

from typing import Optional

import torch

from . import ops


class Rollout:
    def __init__(self, model):
        geo = model.geometry
        if geo.frames != 1 or geo.output_frames != 1:
            raise ValueError("rollout state update needs frames == output_frames == 1 (reference: history_len == 1)")
        self.model = model
        self.n_prog = geo.channels * geo.levels + geo.surface_channels
        self.n_forced = geo.input_only_channels

    def step(self, x: torch.Tensor, forcing: Optional[torch.Tensor] = None, n_dynamic: Optional[int] = None):
        """One forecast step: returns y and updates ``x`` in place for the next step.

        forcing: [B, n_dynamic, 1, H, W] new dynamic-forcing channels (None = carry all forcings).
        """
        y = self.model(x)
        ops.copy_channels(x, y, [(0, 0, self.n_prog)])
        if forcing is not None:
            n_dyn = forcing.shape[1] if n_dynamic is None else n_dynamic
            ops.copy_channels(x, forcing, [(self.n_prog, 0, n_dyn)])
        return y
