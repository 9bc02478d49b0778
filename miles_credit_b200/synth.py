"""Deterministic synthetic weights and inputs (there is no network for checkpoints).

``synthetic_state_dict`` produces a state dict with exactly the reference's keys and shapes
(geometry.state_spec) so it loads with ``strict=True`` into ``credit.models.crossformer.CrossFormer``
and into the B200 module alike.  Spectral-norm ``weight_u/weight_v`` buffers are power-iterated
to convergence, which is what a few train-mode forwards do in the reference
(torch.nn.utils.spectral_norm): without it an ``eval()`` forward of a freshly
built model has outputs of magnitude 1e14.

Each tensor is drawn from its own ``torch.Generator`` seeded by (seed, position of the key), on the
CPU, so the same numbers come out here and on the GPU box.
"""

from __future__ import annotations

import math
from collections import OrderedDict

import torch

from .geometry import Geometry, state_spec


def _gen(seed: int, idx: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + idx * 7919 + 12345) % (2**63 - 1))
    return g


def weight_matrix(w: torch.Tensor, dim: int) -> torch.Tensor:
    """The 2-D view spectral norm works on (dim moved to front, rest flattened)."""
    if dim != 0:
        w = w.permute(dim, *[d for d in range(w.dim()) if d != dim])
    return w.reshape(w.shape[0], -1)


def power_iterate(w_mat: torch.Tensor, g: torch.Generator, iters: int = 30, eps: float = 1e-12):
    u = torch.nn.functional.normalize(torch.randn(w_mat.shape[0], generator=g, dtype=torch.float64), dim=0, eps=eps)
    v = torch.nn.functional.normalize(torch.randn(w_mat.shape[1], generator=g, dtype=torch.float64), dim=0, eps=eps)
    wm = w_mat.double()
    for _ in range(iters):
        v = torch.nn.functional.normalize(wm.t().mv(u), dim=0, eps=eps)
        u = torch.nn.functional.normalize(wm.mv(v), dim=0, eps=eps)
    return u.float(), v.float()


def synthetic_state_dict(geo: Geometry, seed: int = 1000, sn_iters: int = 30) -> "OrderedDict[str, torch.Tensor]":
    return synthesize(state_spec(geo), seed, sn_iters)


def synthesize(spec, seed: int = 1000, sn_iters: int = 30) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic weights for a ``key -> (shape, role)`` table (roles: ``weight:<sn dim>``, ``u``, ``v``, ``bias``,
    ``gain``, ``shift``, ``const:<value>``)."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    keys = list(spec.keys())
    for idx, key in enumerate(keys):
        shape, role = spec[key]
        g = _gen(seed, idx)
        if role.startswith("weight"):
            dim = int(role.split(":")[1])
            fan_in = 1
            for i, s in enumerate(shape):
                if i != dim:
                    fan_in *= s
            if dim != 0:  # ConvTranspose2d: fan-in as torch computes it (shape[1] * k * k)
                fan_in = shape[1] * shape[2] * shape[3]
            bound = 1.0 / math.sqrt(max(fan_in, 1))
            w = (torch.rand(shape, generator=g) * 2 - 1) * bound
            sd[key] = w
            if key.endswith("weight_orig"):
                base = key[: -len("weight_orig")]
                u, v = power_iterate(weight_matrix(w, dim), g, sn_iters)
                sd[base + "weight_u"] = u
                sd[base + "weight_v"] = v
        elif role in ("u", "v"):
            continue  # filled together with weight_orig
        elif role == "bias":
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        elif role == "gain":
            sd[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif role == "shift":
            sd[key] = 0.1 * torch.randn(shape, generator=g)
        elif role.startswith("const:"):
            sd[key] = torch.full(shape, float(role.split(":")[1]))
        else:
            raise AssertionError(role)
    return OrderedDict((k, sd[k]) for k in keys)


def synthetic_input(geo: Geometry, batch: int = 1, seed: int = 1000) -> torch.Tensor:
    """ERA5-shaped z-scored state: N(0,1), the recipe of the reference's own tests (tests/test_models.py:74)."""
    g = _gen(seed, 999983)
    return torch.randn((batch, *geo.in_shape), generator=g, dtype=torch.float32)


def state_checksum(sd) -> float:
    """Order-independent scalar fingerprint used by the golden fixtures to detect generator drift."""
    tot = 0.0
    for k in sorted(sd):
        t = sd[k].double()
        tot += float(t.sum()) + 0.5 * float((t * t).sum())
    return tot
