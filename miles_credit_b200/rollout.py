"""Autoregressive rollout of the forecast step with the state kept on the device.

Mirrors the reference's flat rollout loop (applications/rollout_to_netcdf.py:269-310): ``y = model(x)``,
then ``x <- update_x(x, forcing, y)`` (datasets/gen_2/channel_utils.py:253-291): prognostic channels come
from the prediction, dynamic-forcing channels from the data stream, static channels are carried.
With ``frames > 1`` (FuXi: two input states; CrossFormer configs with a history window) the input slides like the gen2
rollout does (trainers/rollout_utils.py:288-311: drop the oldest time step, append the newest), in one in-place kernel
(``wxf_history_update``); the flat reference loop refuses that case (``build_channel_layout``, channel_utils.py:205-211).
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

    def __init__(self, model, graph: bool = False):
        """``graph=True``: the whole step (kernels, NCCL exchanges, state update) is captured into one CUDA graph on first
        use and replayed afterwards, so the host enqueues one launch per step instead of ~250 (what bounds the decomposed
        step on 8 GPUs, where the average kernel lasts ~15 us).  The returned prediction is then a buffer owned by the
        rollout that the next step overwrites."""
        geo = model.geometry
        if getattr(geo, "output_frames", 1) != 1:
            raise ValueError("rollout state update needs output_frames == 1 (one new state per step)")
        self.frames = geo.frames
        self.model = model
        self.n_prog = geo.channels * geo.levels + geo.surface_channels
        self.n_forced = geo.input_only_channels
        self._y = None  # sharded / graph mode: the full-size prediction buffer is reused
        self.graph = graph
        self._graphs = {}
        self._graphs_version = None
        self.launches_per_replay = 0

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
        if not self.graph:
            return self._step(x, forcing, n_dynamic)
        # captured graphs hold raw pointers into the plan workspaces and the prepared weights: a weight change
        # (load_state_dict, .to(), refresh_weights) bumps the model's version and drops every graph captured before it
        self.model._plan_for(x)
        ver = getattr(self.model, "_weights_version", 0)
        if ver != self._graphs_version:
            self._graphs.clear()
            self._graphs_version = ver
        key = (x.data_ptr(), tuple(x.shape), None if forcing is None else forcing.data_ptr(), n_dynamic)
        entry = self._graphs.get(key)
        if entry is None:
            # plans, NCCL communicators and lazy kernel attributes must exist before the capture: one throw-away step
            self._step(x.clone(), None if forcing is None else forcing, n_dynamic)
            torch.cuda.synchronize(x.device)
            g = torch.cuda.CUDAGraph()
            n0 = ops.LAUNCHES
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                y = self._step(x, forcing, n_dynamic)
            self.launches_per_replay = ops.LAUNCHES - n0
            entry = self._graphs[key] = (g, y)
        g, y = entry
        g.replay()
        ops.LAUNCHES += self.launches_per_replay
        return y

    def _step(self, x: torch.Tensor, forcing: Optional[torch.Tensor], n_dynamic: Optional[int]):
        if not self.sharded:
            if self.graph:  # static output buffer
                xc, plan = self.model._plan_for(x)
                if self._y is None or self._y.shape[0] != x.shape[0]:
                    self._y = torch.empty((x.shape[0], *self.model.geometry.out_shape), device=x.device, dtype=torch.float32)
                with torch.cuda.device(x.device):
                    y = plan.run(xc, self._y)
            else:
                y = self.model(x)
            dev_ctx = torch.cuda.device(x.device) if x.is_cuda else contextlib.nullcontext()
            if self.frames > 1:  # history window: slide the frames, newest from the prediction (+ forcing), one kernel
                n_dyn = 0 if forcing is None else (forcing.shape[1] if n_dynamic is None else n_dynamic)
                with dev_ctx:
                    ops.history_update(x, y, forcing, self.n_prog, n_dyn)
                return y
            with dev_ctx:
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
            if self.frames > 1:
                # history window: slide every frame, newest from the prediction (rows outside this rank's share carry stale
                # values on both sides and are never read by its padding pass)
                n_dyn = 0 if forcing is None else (forcing.shape[1] if n_dynamic is None else n_dynamic)
                with (torch.cuda.device(x.device) if x.is_cuda else contextlib.nullcontext()):
                    ops.history_update(x, y, forcing, self.n_prog, n_dyn)
                return y
            a, b = plan.src_rows[plan.rank]
            x[:, : self.n_prog, :, a:b].copy_(y[:, : self.n_prog, :, a:b])
        if forcing is not None:
            n_dyn = forcing.shape[1] if n_dynamic is None else n_dynamic
            with (torch.cuda.device(x.device) if x.is_cuda else contextlib.nullcontext()):
                ops.copy_channels(x, forcing, [(self.n_prog, 0, n_dyn)])
        return y
