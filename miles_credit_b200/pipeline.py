"""The steps either side of ``y = model(x)`` in the gen2 rollout loop, fused into the boundary passes of
the forecast step instead of running as separate PyTorch modules:

* ``FusedPreblocks``  (f2): ``ERA5Normalizer`` (credit/preblock/norm.py:35-109) + ``ConcatToTensor``
  (credit/preblock/concat.py:37-207).  The per-variable tensors of the batch dict are never concatenated: a table of plane
  addresses drives the padding kernel, which z-scores every value on the way (``wxf_preblock_pad_to_pixel_major``).
* ``FusedPostblocks`` (f3): inverse scaling ``y * std + mean`` (applications/rollout_to_netcdf.py:287; the gen2
  ``bridgescaler`` inverse transform of a standard scaler), ``TracerFixer`` clamps (credit/postblock/conservation.py:88-115)
  — both in the epilogue of the un-pad / resize / NCHW pass — and ``GlobalMassFixer`` (conservation.py:118-176, hybrid-sigma
  midpoint grid): one column-integral reduction kernel + one rescale of the surface pressure.
  ``Reconstruct`` (credit/postblock/reconstruct.py:62-84) is a view: ``split()`` returns the per-variable views.
* ``ForecastHandoff`` (f4): pinned, double-buffered device -> host hand-off of the prediction for the writer pool
  (``save_output_fn`` in credit/trainers/rollout_utils.py:286; ``ForecastWriter`` in credit/output_gen2.py), overlapped with
  the next step on a copy stream.

Host logic only; every tensor operation is a kernel of the C ABI (no PyTorch fallback).
"""

from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops

# canonical cross-group concat rank (credit/datasets/gen_2/channel_utils.py:88-93)
FIELD_TYPE_RANK = {"prognostic": 0, "static": 1, "dynamic_forcing": 2, "diagnostic": 3}
GRAVITY = 9.80665  # credit/physics_constants.py


def channel_sort_key(var_key: str) -> tuple:
    """``_channel_sort_key`` of credit/preblock/concat.py:24-34: (field-type rank, 3d before 2d); stable within a bucket."""
    parts = var_key.split("/")
    ft = parts[1] if len(parts) > 1 else ""
    dim = parts[2] if len(parts) > 2 else ""
    return (FIELD_TYPE_RANK.get(ft, len(FIELD_TYPE_RANK)), 0 if dim == "3d" else 1)


def input_channel_map(batch_input: Dict[str, Dict[str, torch.Tensor]]):
    """``metadata["input"]["_channel_map"]`` as ConcatToTensor builds it (concat.py:124-137): var_key -> slice, orig_shape."""
    out = OrderedDict()
    cursor = 0
    for _source, variables in batch_input.items():
        for key, t in sorted(variables.items(), key=lambda kv: channel_sort_key(kv[0])):
            n_levels, T = int(t.shape[1]), int(t.shape[2])
            out[key] = {"slice": slice(cursor, cursor + n_levels * T), "orig_shape": (n_levels, T)}
            cursor += n_levels * T
    return out


class FusedPreblocks:
    """Normalisation statistics + the channel order, turned into the tables the fused padding kernel reads.

    ``mean`` / ``std``: ``{varname: tensor}`` — scalar for 2-D variables, a 1-D level vector for 3-D ones, exactly what
    ``ERA5Normalizer.__init__`` extracts from its NetCDF files (norm.py:58-72); variables without statistics pass through
    unchanged (norm.py:86-87)."""

    def __init__(self, mean: Dict[str, torch.Tensor], std: Dict[str, torch.Tensor], levels: Optional[Sequence[int]] = None):
        idx = [lv - 1 for lv in levels] if levels is not None else None
        self.mean, self.std = {}, {}
        for var in set(mean) & set(std):
            m, s = torch.as_tensor(mean[var], dtype=torch.float32), torch.as_tensor(std[var], dtype=torch.float32)
            if idx is not None and m.dim() == 1 and m.shape[0] > 1:
                m, s = m[idx], s[idx]
            self.mean[var], self.std[var] = m, s
        self._cache = {}

    def tables(self, batch_input: Dict[str, Dict[str, torch.Tensor]]):
        """(plane-address table int64 [B*C], mean [C], std [C], B, C, T, H, W, keep-alive list) for a batch ``input`` dict of
        [B, n_levels, T, H, W] fp32 CUDA tensors.  Cached on the tensors' addresses (a rollout re-uses its buffers)."""
        tensors: List[Tuple[str, torch.Tensor]] = []
        for _source, variables in batch_input.items():
            tensors += sorted(variables.items(), key=lambda kv: channel_sort_key(kv[0]))
        if not tensors:
            raise ValueError("No 'input' tensors found in batch.")
        sig = tuple((k, t.data_ptr(), tuple(t.shape)) for k, t in tensors)
        hit = self._cache.get(sig)
        if hit is not None:
            return hit
        B, _, T, H, W = tensors[0][1].shape
        dev = tensors[0][1].device
        addr = [[] for _ in range(B)]
        means, stds = [], []
        for key, t in tensors:
            if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError(f"{key}: the fused pre-blocks take contiguous fp32 CUDA tensors (no CPU path)")
            if tuple(t.shape[2:]) != (T, H, W) or t.shape[0] != B:
                raise ValueError(f"{key}: shape {tuple(t.shape)} does not match [B={B}, levels, T={T}, {H}, {W}]")
            n_lev = int(t.shape[1])
            var = key.split("/")[-1]
            if var in self.mean:
                m, s = self.mean[var], self.std[var]
                per_level = m.dim() == 1 and m.shape[0] > 1
                if per_level and m.shape[0] != n_lev:
                    raise ValueError(f"{key}: {n_lev} levels but statistics for {m.shape[0]}")
                means += [float(m[l]) if per_level else float(m.reshape(-1)[0]) for l in range(n_lev)]
                stds += [float(s[l]) if per_level else float(s.reshape(-1)[0]) for l in range(n_lev)]
            else:
                means += [0.0] * n_lev
                stds += [1.0] * n_lev
            plane = T * H * W * 4
            for b in range(B):
                base = t.data_ptr() + b * n_lev * plane
                addr[b] += [base + l * plane for l in range(n_lev)]
        C = len(means)
        table = torch.tensor([a for row in addr for a in row], dtype=torch.int64, device=dev)
        out = (table, torch.tensor(means, dtype=torch.float32, device=dev), torch.tensor(stds, dtype=torch.float32, device=dev),
               int(B), C, int(T), int(H), int(W), [t for _, t in tensors])
        self._cache[sig] = out
        return out

    def materialise(self, batch_input) -> torch.Tensor:
        """The tensor ``ConcatToTensor`` would return, [B, C, T, H, W], produced by the same kernel with zero padding (a check /
        convenience path: the forecast step never needs it)."""
        table, mean, std, B, C, T, H, W, _keep = self.tables(batch_input)
        ld = (C * T + 7) // 8 * 8
        pm = torch.empty((B, H, W, ld), device=table.device, dtype=torch.float32)
        ops.preblock_pad_to_pixel_major(table, mean, std, B, C, T, H, W, (0, 0), (0, 0), "earth", ld, out=pm)
        return pm[..., : C * T].permute(0, 3, 1, 2).reshape(B, C, T, H, W).contiguous()


class FusedPostblocks:
    """Per-output-channel inverse scaling and tracer clamps for the epilogue of the un-pad pass, plus the global dry-air
    mass fixer.  ``channel_map``: ``metadata["target"]["_channel_map"]`` (var_key -> slice / orig_shape)."""

    def __init__(self, channel_map, n_channels: int, mean: Optional[Dict[str, torch.Tensor]] = None,
                 std: Optional[Dict[str, torch.Tensor]] = None, tracer_vars: Sequence[str] = (), tracer_thres=0.0,
                 tracer_thres_max=None, device="cuda"):
        self.channel_map = channel_map
        scale = torch.ones(n_channels)
        shift = torch.zeros(n_channels)
        lo = torch.full((n_channels,), float("-inf"))
        hi = torch.full((n_channels,), float("inf"))
        for key, info in channel_map.items():
            sl, var = info["slice"], key.split("/")[-1]
            if mean is not None and var in mean:
                m, s = torch.as_tensor(mean[var], dtype=torch.float32).reshape(-1), torch.as_tensor(std[var], dtype=torch.float32).reshape(-1)
                n = sl.stop - sl.start
                scale[sl] = s if s.numel() == n else s.expand(n)
                shift[sl] = m if m.numel() == n else m.expand(n)
        n_t = len(tracer_vars)
        los = list(tracer_thres) if isinstance(tracer_thres, (list, tuple)) else [tracer_thres] * n_t
        his = (list(tracer_thres_max) if isinstance(tracer_thres_max, (list, tuple)) else [tracer_thres_max] * n_t)
        for key, a, b in zip(tracer_vars, los, his):
            sl = channel_map[key]["slice"]
            lo[sl] = float(a)
            if b is not None:
                hi[sl] = float(b)
        self.scale, self.shift, self.lo, self.hi = (t.to(device).contiguous() for t in (scale, shift, lo, hi))

    def split(self, y: torch.Tensor):
        """``Reconstruct.forward`` (reconstruct.py:62-84): nested dict of [B, n_levels, n_time, H, W] views of y."""
        flat = y.flatten(1, 2) if y.dim() == 5 else y
        out = {}
        for key, info in self.channel_map.items():
            out.setdefault(key.split("/")[0], {})[key] = flat[:, info["slice"]].unflatten(1, tuple(info["orig_shape"]))
        return out


class GlobalMassFixerB200:
    """``GlobalMassFixer`` (credit/postblock/conservation.py:118-176) on the hybrid-sigma grid with midpoint quantities: the
    dry-air mass of the input state sets the target, the predicted surface pressure is rescaled to match.

    sums(state) -> (A, B) with A = sum_pixels area * sum_l da_l (1 - q_l),  B = sum_pixels area * sp * sum_l db_l (1 - q_l):
    mass_t0 = (A0 + B0) / g from the input (physics_core.total_dry_air_mass, :500-508), mass_a = A1 / g, mass_b = B1 / g
    from the prediction, ratio = (mass_t0 - mass_a) / mass_b, sp *= ratio."""

    def __init__(self, area: torch.Tensor, coef_a: torch.Tensor, coef_b: torch.Tensor, device="cuda"):
        self.area = area.to(device=device, dtype=torch.float32).contiguous()
        self.da = coef_a.diff().to(device=device, dtype=torch.float32).contiguous()
        self.db = coef_b.diff().to(device=device, dtype=torch.float32).contiguous()
        self.n_levels = int(self.da.numel())

    def sums(self, q: torch.Tensor, sp: torch.Tensor, rows=None) -> torch.Tensor:
        """q: [B, L, H, W] view (level stride = plane stride), sp: [B, H, W] view -> fp64 [B, 2]; ``rows`` = (first, count)
        restricts the sum to a latitude band (a domain-decomposed caller all-reduces the partial sums)."""
        return ops.dry_mass_sums(q, sp, self.area, self.da, self.db, rows)

    def apply(self, q_pred, sp_pred, q_input, sp_input):
        """Rescales ``sp_pred`` in place; returns the ratio [B] (fp32, computed like the reference: fp32 division of sums)."""
        s0, s1 = self.sums(q_input, sp_input), self.sums(q_pred, sp_pred)
        mass_t0 = ((s0[:, 0] + s0[:, 1]) / GRAVITY).float()
        ratio = ((mass_t0 - (s1[:, 0] / GRAVITY).float()) / (s1[:, 1] / GRAVITY).float()).contiguous()
        ops.scale_planes(sp_pred, ratio)
        return ratio


class GlobalWaterFixerB200:
    """``GlobalWaterFixer`` (credit/postblock/conservation.py:179-236), hybrid-sigma grid with midpoint quantities: the
    column-water tendency plus evaporation must balance precipitation; precipitation is rescaled by
    ratio = (-sum dTWC/dt - sum E) / sum P (area-weighted global sums, one kernel)."""

    def __init__(self, area: torch.Tensor, coef_a: torch.Tensor, coef_b: torch.Tensor, lead_time_periods: int, device="cuda"):
        f32 = dict(device=device, dtype=torch.float32)
        self.area, self.coef_a, self.coef_b = (t.to(**f32).contiguous() for t in (area, coef_a, coef_b))
        self.n_seconds = int(lead_time_periods) * 3600

    def sums(self, q_pred, sp_pred, q_input, sp_input, precip, evapor, rows=None) -> torch.Tensor:
        """fp64 [B, 3]; ``rows`` = (first, count) restricts the sums to a latitude band (a decomposed caller adds the bands)."""
        return ops.water_budget_sums(q_pred, sp_pred, q_input, sp_input, precip, evapor, self.area, self.coef_a, self.coef_b,
                                     self.n_seconds, rows)

    def apply(self, q_pred, sp_pred, q_input, sp_input, precip, evapor):
        """Rescales ``precip`` ([B, H, W] view of the prediction) in place; returns the ratio [B] (fp32 like the reference)."""
        s = self.sums(q_pred, sp_pred, q_input, sp_input, precip, evapor).float()
        twc, e, p = s[:, 0], s[:, 1], s[:, 2]
        residual = -twc - e - p
        ratio = ((p + residual) / p).contiguous()
        ops.scale_planes(precip, ratio)
        return ratio


class GlobalEnergyFixerB200:
    """``GlobalEnergyFixerUpDown`` (credit/postblock/conservation.py:239-376), hybrid-sigma grid with midpoint quantities: the
    column total-energy tendency is forced to match the net TOA + surface fluxes, temperature carries the correction.  One
    kernel forms the four global sums (R_T, F_S, TE(t0), TE(t1)), a second one rewrites T in place."""

    def __init__(self, area, coef_a, coef_b, gph_surf, lead_time_periods: int, device="cuda"):
        f32 = dict(device=device, dtype=torch.float32)
        self.area, self.coef_a, self.coef_b, self.gph_surf = (t.to(**f32).contiguous() for t in (area, coef_a, coef_b, gph_surf))
        self.n_seconds = int(lead_time_periods) * 3600

    def describe(self, pred3, pred2, in3, sp_input, toa_down_input, rows=None):
        return ops.make_energy_desc(pred3, pred2, in3, sp_input, toa_down_input, self.gph_surf, self.area, self.coef_a,
                                    self.coef_b, self.n_seconds, rows)

    def apply(self, pred3, pred2, in3, sp_input, toa_down_input):
        """pred3 = (T, q, U, V), pred2 = (sp, toa_up_solar, toa_up_olr, surf_down_solar, surf_up_solar, surf_down_lw, surf_up_lw,
        surf_sh, surf_lh) views of the prediction, in3 = (T, q, U, V) of the last input frame.  T is corrected in place;
        returns the ratio [B]."""
        desc = self.describe(pred3, pred2, in3, sp_input, toa_down_input)
        s = ops.energy_budget_sums(desc, pred3[0].device).float()
        r_t, f_s, te0, te1 = s[:, 0], s[:, 1], s[:, 2], s[:, 3]
        ratio = ((self.n_seconds * (r_t - f_s) + te0) / te1).contiguous()
        ops.energy_fix_temperature(desc, ratio)
        return ratio


class ForecastHandoff:
    """Device -> pinned-host hand-off of every step's prediction, double buffered on a copy stream so the D2H of step k
    overlaps the forward of step k + 1.  ``push(y)`` returns at once; ``pop()`` yields (step, host tensor) in order once that
    copy has landed — the tensor is what the reference hands to its writer pool (``y_processed`` ->
    ``save_output_fn``, trainers/rollout_utils.py:286).  A host buffer is recycled ``depth`` pushes later."""

    def __init__(self, shape, device, depth: int = 2, rows: Optional[Tuple[int, int]] = None):
        self.rows = rows  # (first, last) of the latitude rows this rank owns (decomposed forecast): only those are copied
        shp = list(shape)
        if rows is not None:
            shp[-2] = rows[1] - rows[0]
        self.host = [torch.empty(shp, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.stage = [torch.empty(shp, dtype=torch.float32, device=device) for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self.busy = [False] * depth
        self.stream = torch.cuda.Stream(device=device)
        self.step = 0
        self.pending: List[Tuple[int, int]] = []
        self.bytes_per_step = self.host[0].numel() * 4

    def push(self, y: torch.Tensor):
        i = self.step % len(self.host)
        if self.busy[i]:
            self.done[i].synchronize()  # the host buffer is still owned by a copy `depth` steps old
        src = y if self.rows is None else y[..., self.rows[0]: self.rows[1], :]
        self.stage[i].copy_(src)        # the rollout overwrites its prediction buffer next step: snapshot on the device
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            self.host[i].copy_(self.stage[i], non_blocking=True)
            self.done[i].record(self.stream)
        self.busy[i] = True
        self.pending.append((self.step, i))
        self.step += 1

    def pop(self, block: bool = True):
        if not self.pending:
            return None
        step, i = self.pending[0]
        if not block and not self.done[i].query():
            return None
        self.done[i].synchronize()
        self.pending.pop(0)
        return step, self.host[i]

    def drain(self):
        out = []
        while self.pending:
            out.append(self.pop())
        return out
