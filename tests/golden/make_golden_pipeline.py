"""Golden vectors for the pre / post-blocks either side of the forecast step, from the UNMODIFIED
reference classes (run in the build container only):

    python tests/golden/make_golden_pipeline.py

* ``ERA5Normalizer`` (credit/preblock/norm.py) + ``ConcatToTensor`` (credit/preblock/concat.py) on a small batch dict;
* ``Reconstruct`` (credit/postblock/reconstruct.py), the gen1 inverse scaling ``y * std + mean``
  (applications/rollout_to_netcdf.py:287), ``TracerFixer`` and ``GlobalMassFixer`` (credit/postblock/conservation.py) with
  the reference's ``physics_hybrid_sigma_level`` core (credit/physics_core.py).

The reference modules import xarray / credit.data (absent here); stubs are placed in ``sys.modules`` first and the objects
that would read NetCDF files in ``__init__`` are created with ``__new__`` and given their statistics directly — the
arithmetic that runs is the reference's own ``forward``.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.modules["xarray"] = types.ModuleType("xarray")
data_stub = types.ModuleType("credit.data")
data_stub.get_forward_data = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stubbed"))
sys.modules["credit.data"] = data_stub

from credit.physics_core import physics_hybrid_sigma_level  # noqa: E402
from credit.postblock.conservation import GlobalMassFixer, TracerFixer  # noqa: E402
from credit.postblock.reconstruct import Reconstruct  # noqa: E402
from credit.preblock.concat import ConcatToTensor  # noqa: E402
from credit.preblock.norm import ERA5Normalizer  # noqa: E402


def main():
    g = torch.Generator().manual_seed(20260117)
    B, L, T, H, W = 2, 5, 1, 12, 20
    rnd = lambda *s: torch.randn(*s, generator=g)  # noqa: E731

    # ---- pre-blocks ---------------------------------------------------------------------------------------------
    # insertion order deliberately differs from the canonical channel order (2d before 3d, forcing before static)
    inp = {"era5": {
        "era5/prognostic/2d/SP": 1e5 + 500 * rnd(B, 1, T, H, W),
        "era5/dynamic_forcing/2d/tsi": rnd(B, 1, T, H, W).abs() * 1000,
        "era5/prognostic/3d/T": 250 + 20 * rnd(B, L, T, H, W),
        "era5/prognostic/3d/Q": (0.004 + 0.002 * rnd(B, L, T, H, W)).abs(),
        "era5/static/2d/Z_GDS4_SFC": rnd(B, 1, T, H, W) * 3000,
        "era5/static/2d/LSM": torch.rand(B, 1, T, H, W, generator=g),     # no statistics: passes through
        "era5/prognostic/2d/t2m": 280 + 15 * rnd(B, 1, T, H, W),
    }}
    mean = {"T": 250 + rnd(L), "Q": (0.004 + 0.001 * rnd(L)).abs(), "SP": torch.tensor(1e5), "t2m": torch.tensor(281.0),
            "tsi": torch.tensor(400.0), "Z_GDS4_SFC": torch.tensor(120.0)}
    std = {"T": 15 + rnd(L).abs(), "Q": torch.tensor([2e-3, 1e-3, 0.0, 5e-4, 1e-3]),   # one zero std: clamped to 1e-12
           "SP": torch.tensor(900.0), "t2m": torch.tensor(14.0), "tsi": torch.tensor(350.0), "Z_GDS4_SFC": torch.tensor(2500.0)}
    norm = ERA5Normalizer.__new__(ERA5Normalizer)
    torch.nn.Module.__init__(norm)
    norm._mean, norm._std = dict(mean), dict(std)
    batch = {"input": inp}
    x_ref, meta = ConcatToTensor()(norm(batch))
    cmap = {k: (v["slice"].start, v["slice"].stop, tuple(v["orig_shape"])) for k, v in meta["input"]["_channel_map"].items()}

    # ---- post-blocks ----------------------------------------------------------------------------------------------
    tmap = {
        "era5/prognostic/3d/T": {"slice": slice(0, L), "orig_shape": (L, 1)},
        "era5/prognostic/3d/Q": {"slice": slice(L, 2 * L), "orig_shape": (L, 1)},
        "era5/prognostic/2d/SP": {"slice": slice(2 * L, 2 * L + 1), "orig_shape": (1, 1)},
        "era5/prognostic/2d/t2m": {"slice": slice(2 * L + 1, 2 * L + 2), "orig_shape": (1, 1)},
        "era5/diagnostic/2d/tp": {"slice": slice(2 * L + 2, 2 * L + 3), "orig_shape": (1, 1)},
    }
    C_out = 2 * L + 3
    y_pred = rnd(B, C_out, 1, H, W)                       # normalised model output
    omean = dict(mean, tp=torch.tensor(0.002))
    ostd = dict(std, tp=torch.tensor(0.004))
    ostd["Q"] = torch.tensor([2e-3, 1e-3, 1.5e-3, 5e-4, 1e-3])
    bd = Reconstruct()({"y_pred": y_pred, "metadata": {"target": {"_channel_map": tmap}}})
    # inverse scaling, per variable: y * std + mean  (rollout_to_netcdf.py:287)
    for key, t in list(bd["y_processed"]["era5"].items()):
        var = key.split("/")[-1]
        m, s = omean[var], ostd[var]
        if m.dim() == 1:
            m, s = m.view(1, -1, 1, 1, 1), s.view(1, -1, 1, 1, 1)
        bd["y_processed"]["era5"][key] = t * s + m
    bd = TracerFixer(["era5/prognostic/3d/Q", "era5/diagnostic/2d/tp"], [1e-9, 0.0], [None, 0.01])(bd)
    scaled = {k: v.clone() for k, v in bd["y_processed"]["era5"].items()}

    lat = torch.linspace(87.0, -87.0, H)
    lon = torch.linspace(0.0, 360.0 - 360.0 / W, W)
    lon2d, lat2d = torch.meshgrid(lon, lat, indexing="xy")
    coef_a = torch.tensor([0.0, 2000.0, 6000.0, 9000.0, 4000.0, 0.0])      # L + 1 interfaces (midpoint quantities)
    coef_b = torch.tensor([0.0, 0.0, 0.05, 0.3, 0.75, 1.0])
    fixer = GlobalMassFixer.__new__(GlobalMassFixer)
    torch.nn.Module.__init__(fixer)
    fixer.q_var, fixer.sp_var, fixer.input_source_key = "era5/prognostic/3d/Q", "era5/prognostic/2d/SP", "x_physical"
    fixer.core = physics_hybrid_sigma_level(lon2d, lat2d, coef_a, coef_b, midpoint=True)
    fixer.flag_sigma, fixer.midpoint, fixer.N_levels, fixer.coef_a, fixer.coef_b = True, True, L, coef_a, coef_b
    bd["x_physical"] = {"era5": {k: v for k, v in inp["era5"].items()}}
    bd = fixer(bd)
    fixed_sp = bd["y_processed"]["era5"]["era5/prognostic/2d/SP"]

    torch.save({
        "input": inp["era5"], "mean": mean, "std": std, "x_ref": x_ref, "channel_map": cmap,
        "y_pred": y_pred, "target_map": {k: (v["slice"].start, v["slice"].stop, tuple(v["orig_shape"])) for k, v in tmap.items()},
        "out_mean": omean, "out_std": ostd, "tracer": (["era5/prognostic/3d/Q", "era5/diagnostic/2d/tp"], [1e-9, 0.0], [None, 0.01]),
        "scaled": scaled, "area": fixer.core.area, "coef_a": coef_a, "coef_b": coef_b, "fixed_sp": fixed_sp,
    }, os.path.join(HERE, "pipeline.pt"))
    print("x_ref", tuple(x_ref.shape), "channels", list(cmap), "sp ratio", (fixed_sp / scaled["era5/prognostic/2d/SP"]).flatten()[:2])


if __name__ == "__main__":
    main()
