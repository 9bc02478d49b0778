"""Golden vectors for the remaining global conservation fixers of SURVEY.md section 8 f3, from the UNMODIFIED reference
classes ``GlobalWaterFixer`` and ``GlobalEnergyFixerUpDown`` (credit/postblock/conservation.py:179-236, 239-376) with the
reference's ``physics_hybrid_sigma_level`` core (credit/physics_core.py).  Run in the build container only:

    python tests/golden/make_golden_fixers.py

xarray / credit.data are absent here: stubs go into ``sys.modules`` and the fixers are created with ``__new__`` and given the
attributes their ``__init__`` would read from the physics NetCDF file — the arithmetic that runs is the reference's ``forward``.
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
from credit.postblock.conservation import GlobalEnergyFixerUpDown, GlobalWaterFixer  # noqa: E402

P3, P2, D2, F2 = "era5/prognostic/3d/", "era5/prognostic/2d/", "era5/diagnostic/2d/", "era5/dynamic_forcing/2d/"


def main():
    g = torch.Generator().manual_seed(20260118)
    B, L, T, H, W = 2, 6, 2, 10, 16
    rnd = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    lat = torch.linspace(85.0, -85.0, H)
    lon = torch.linspace(0.0, 360.0 - 360.0 / W, W)
    lon2d, lat2d = torch.meshgrid(lon, lat, indexing="xy")
    coef_a = torch.tensor([0.0, 1500.0, 5000.0, 9000.0, 7000.0, 3000.0, 0.0])     # L + 1 interfaces (midpoint quantities)
    coef_b = torch.tensor([0.0, 0.0, 0.02, 0.15, 0.45, 0.8, 1.0])
    core = physics_hybrid_sigma_level(lon2d, lat2d, coef_a, coef_b, midpoint=True)

    def state(t_frames):
        return {
            P3 + "T": 250 + 25 * rnd(B, L, t_frames, H, W),
            P3 + "Q": (0.004 + 0.003 * rnd(B, L, t_frames, H, W)).abs(),
            P3 + "U": 12 * rnd(B, L, t_frames, H, W),
            P3 + "V": 9 * rnd(B, L, t_frames, H, W),
            P2 + "SP": 1e5 + 1500 * rnd(B, 1, t_frames, H, W),
        }

    x_phys = state(T)                                                   # two input frames: the fixers read the last one
    x_phys[F2 + "SOLIN"] = 1361.0 * torch.rand(B, 1, T, H, W, generator=g)
    y = state(1)                                                        # the prediction, physical units
    for k, scale in (("tp", 0.004), ("evap", 0.002)):                   # accumulated over the step [m]
        y[D2 + k] = scale * torch.rand(B, 1, 1, H, W, generator=g) * (1.0 if k == "tp" else -1.0)
    for k, scale in (("toa_up_sw", 100.0), ("toa_up_lw", 240.0)):        # TOA fluxes: mean W / m^2 (the fixer multiplies by N)
        y[D2 + k] = scale * (0.5 + torch.rand(B, 1, 1, H, W, generator=g))
    for k, scale in (("sfc_dn_sw", 180.0), ("sfc_up_sw", 30.0), ("sfc_dn_lw", 340.0), ("sfc_up_lw", 390.0), ("sfc_sh", -20.0),
                     ("sfc_lh", -80.0)):                                 # surface fluxes: accumulated J / m^2 over the 6 h step
        y[D2 + k] = scale * (0.5 + torch.rand(B, 1, 1, H, W, generator=g)) * 21600.0
    gph_surf = 3000.0 * torch.rand(H, W, generator=g)

    def fresh():
        return {"y_processed": {"era5": {k: v.clone() for k, v in y.items()}}, "x_physical": {"era5": dict(x_phys)}}

    def physics(obj):
        obj.core, obj.flag_sigma, obj.midpoint, obj.N_levels, obj.coef_a, obj.coef_b = core, True, True, L, coef_a, coef_b
        obj.input_source_key = "x_physical"

    wf = GlobalWaterFixer.__new__(GlobalWaterFixer)
    torch.nn.Module.__init__(wf)
    wf.q_var, wf.sp_var, wf.precip_var, wf.evapor_var, wf.N_seconds = P3 + "Q", P2 + "SP", D2 + "tp", D2 + "evap", 6 * 3600
    physics(wf)
    out_w = wf(fresh())["y_processed"]["era5"][D2 + "tp"]

    ef = GlobalEnergyFixerUpDown.__new__(GlobalEnergyFixerUpDown)
    torch.nn.Module.__init__(ef)
    ef.T_var, ef.q_var, ef.U_var, ef.V_var, ef.sp_var = P3 + "T", P3 + "Q", P3 + "U", P3 + "V", P2 + "SP"
    ef.toa_down_solar_input_var = F2 + "SOLIN"
    ef.toa_up_solar_var, ef.toa_up_olr_var = D2 + "toa_up_sw", D2 + "toa_up_lw"
    ef.surf_down_solar_var, ef.surf_up_solar_var = D2 + "sfc_dn_sw", D2 + "sfc_up_sw"
    ef.surf_down_lw_var, ef.surf_up_lw_var = D2 + "sfc_dn_lw", D2 + "sfc_up_lw"
    ef.surf_sh_var, ef.surf_lh_var = D2 + "sfc_sh", D2 + "sfc_lh"
    ef.N_seconds, ef.GPH_surf = 6 * 3600, gph_surf
    physics(ef)
    out_e = ef(fresh())["y_processed"]["era5"][P3 + "T"]

    torch.save({"x_physical": x_phys, "y": y, "gph_surf": gph_surf, "area": core.area, "coef_a": coef_a, "coef_b": coef_b,
                "n_seconds": 6 * 3600, "water_fixed_tp": out_w, "energy_fixed_T": out_e,
                "names": {"T": P3 + "T", "Q": P3 + "Q", "U": P3 + "U", "V": P3 + "V", "SP": P2 + "SP", "SOLIN": F2 + "SOLIN",
                          "tp": D2 + "tp", "evap": D2 + "evap", "toa_up_sw": D2 + "toa_up_sw", "toa_up_lw": D2 + "toa_up_lw",
                          "sfc_dn_sw": D2 + "sfc_dn_sw", "sfc_up_sw": D2 + "sfc_up_sw", "sfc_dn_lw": D2 + "sfc_dn_lw",
                          "sfc_up_lw": D2 + "sfc_up_lw", "sfc_sh": D2 + "sfc_sh", "sfc_lh": D2 + "sfc_lh"}},
               os.path.join(HERE, "fixers.pt"))
    print("precip ratio", (out_w / y[D2 + "tp"]).flatten()[:2].tolist(), "T change max", float((out_e - y[P3 + "T"]).abs().max()))


if __name__ == "__main__":
    main()
