"""Autoregressive rollout of the forecast step with the state kept on the device.

Mirrors the reference's flat rollout loop (applications/rollout_to_netcdf.py:269-310): ``y = model(x)``,
then ``x <- update_x(x, forcing, y)`` (datasets/gen_2/channel_utils.py:253-291): prognostic channels come
from the prediction, dynamic-forcing channels from the data stream, static channels are carried.
Only single-frame inputs (history_len == 1) are supported, like ``build_channel_layout`` (channel_utils.py:205-211).
"""

from __future__ import annotations

import contextlib
from typing import Optional

import torch

from . import ops


class Rollout:
    """``step(x)`` advances the state in place and returns the prediction.

    With a domain-decomposed model (``domain.convert_to_domain_parallel``) the state stays SHARDED between steps: every
    rank passes a full-size ``x`` of which only the rows its next padding pass reads are kept current
    (``plan.src_rows[rank]``), and the returned prediction holds this rank's rows (``plan.out_rows[rank]``; the other rows
    are stale).  Per step the ranks exchange just the halo rows of the new prognostic channels, not the whole field
    (the reference all-gathers every field between steps, credit/trainers/trainer_gen2.py:260-266).
    ``gather(y)`` assembles the full prediction on every rank when a caller needs it.
    """

    def __init__(self, model):
        geo = model.geometry
        if geo.frames != 1 or geo.output_frames != 1:
            raise ValueError("rollout state update needs frames == output_frames == 1 (reference: history_len == 1)")
        self.model = model
        self.n_prog = geo.channels * geo.levels + geo.surface_channels
        self.n_forced = geo.input_only_channels
        self._y = None  # sharded mode: the full-size prediction buffer is reused (only this rank's rows are rewritten)

    @property
    def sharded(self) -> bool:
        return getattr(self.model, "_domain", None) is not None

    def own_rows(self, x: torch.Tensor):
        """Rows of the prediction this rank computes (everything when the model is not decomposed)."""
        if not self.sharded:
            return 0, self.model.geometry.h_out
        _, plan = self.model._plan_for(x)
        return plan.out_rows[plan.rank]

    def gather(self, y: torch.Tensor) -> torch.Tensor:
        """Sharded mode: fill the other ranks' rows of ``y`` (collective; every rank of the domain group must call it)."""
        if self.sharded:
            plan = next(iter(self.model._plans.values()))
            plan.exchange_rows(y, y.shape[1], [(0, y.shape[-2])] * plan.world)
        return y

    @torch.no_grad()
    def step(self, x: torch.Tensor, forcing: Optional[torch.Tensor] = None, n_dynamic: Optional[int] = None):
        """One forecast step: returns y and updates ``x`` in place for the next step.

        forcing: [B, n_dynamic, 1, H, W] new dynamic-forcing channels (None = carry all forcings).
        """
        if not self.sharded:
            y = self.model(x)
            ops.copy_channels(x, y, [(0, 0, self.n_prog)])
        else:
            xc, plan = self.model._plan_for(x)
            if xc.data_ptr() != x.data_ptr():
                raise ValueError("sharded rollout updates the state in place: pass a contiguous fp32 tensor")
            if self._y is None:
                self._y = torch.zeros((1, *self.model.geometry.out_shape), device=x.device, dtype=torch.float32)
            with (torch.cuda.device(x.device) if x.is_cuda else contextlib.nullcontext()):
                y = plan.run_band(x, self._y)
                # rows of the new state the next padding pass of each rank reads: own rows + halo rows from the neighbours
                plan.exchange_rows(y, self.n_prog, plan.src_rows)
            a, b = plan.src_rows[plan.rank]
            x[:, : self.n_prog, :, a:b].copy_(y[:, : self.n_prog, :, a:b])
        if forcing is not None:
            n_dyn = forcing.shape[1] if n_dynamic is None else n_dynamic
            ops.copy_channels(x, forcing, [(self.n_prog, 0, n_dyn)])
        return y
