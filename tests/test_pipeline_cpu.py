"""CPU: host logic of the fused pre / post-blocks (miles_credit_b200/pipeline.py) through the C-ABI emulator, against the
golden vectors of the UNMODIFIED reference classes (tests/golden/make_golden_pipeline.py: ERA5Normalizer, ConcatToTensor,
Reconstruct, TracerFixer)."""
import os

import pytest
import torch

from miles_credit_b200 import lib as wlib
from miles_credit_b200 import ops, pipeline

from abi_emulator import EmulatedLib


@pytest.fixture
def emulated(monkeypatch):
    emu = EmulatedLib()
    monkeypatch.setattr(wlib, "_lib", emu)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(ops, "_req", lambda *a, **k: None)
    return emu


@pytest.fixture(scope="module")
def fx(golden_dir):
    return torch.load(os.path.join(golden_dir, "pipeline.pt"), weights_only=False)


def test_channel_order_is_the_reference_concat_order(fx):
    cmap = pipeline.input_channel_map({"era5": fx["input"]})
    assert list(cmap) == list(fx["channel_map"])                      # prognostic 3d, 2d, static, dynamic forcing
    for k, (a, b, shp) in fx["channel_map"].items():
        assert (cmap[k]["slice"].start, cmap[k]["slice"].stop, tuple(cmap[k]["orig_shape"])) == (a, b, shp)
    assert pipeline.channel_sort_key("era5/static/2d/LSM") < pipeline.channel_sort_key("era5/dynamic_forcing/2d/tsi")


def test_fused_preblocks_tables_and_emulated_kernel(fx, emulated, monkeypatch):
    pre = pipeline.FusedPreblocks(fx["mean"], fx["std"])
    inp = {"era5": {k: v.contiguous() for k, v in fx["input"].items()}}
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))   # host logic only: pretend-device tensors
    table, mean, std, B, C, T, H, W, keep = pre.tables(inp)
    assert (B, C, T, H, W) == tuple(fx["x_ref"].shape) and table.numel() == B * C
    assert float(std[7]) == 0.0 and float(mean[13]) == 0.0 and float(std[13]) == 1.0   # zero std kept (kernel clamps); LSM passes
    assert pre.tables(inp) is pre.tables(inp)                                           # cached on the buffer addresses
    x = pre.materialise(inp)
    assert torch.equal(x, fx["x_ref"])                                                  # bit-exact vs ERA5Normalizer + ConcatToTensor


def test_fused_postblocks_tables_and_emulated_epilogue(fx, emulated):
    tmap = {k: {"slice": slice(a, b), "orig_shape": shp} for k, (a, b, shp) in fx["target_map"].items()}
    names, los, his = fx["tracer"]
    C = fx["y_pred"].shape[1]
    post = pipeline.FusedPostblocks(tmap, C, fx["out_mean"], fx["out_std"], names, los, his, device="cpu")
    y = fx["y_pred"]
    B, _, _, H, W = y.shape
    pm = y[:, :, 0].permute(0, 2, 3, 1).contiguous()                                   # pixel-major decoder output, ld = C
    out = torch.empty(B, C, 1, H, W)
    ops.unpad_resize_post_to_nchw(pm, C, out, B, C, H, W, 0, 0, H, W, H, W, post.scale, post.shift, post.lo, post.hi)
    got = post.split(out)["era5"]
    for k, ref in fx["scaled"].items():
        assert got[k].shape == ref.shape
        assert torch.equal(got[k], ref), k                                              # Reconstruct + y*std+mean + TracerFixer


def _fixer_views(fx, dev):
    """The prediction packed into ONE [B, C, 1, H, W] tensor (the model's output layout) and channel views of it; the input
    state as one tensor per variable (the batch dict's layout), last frame."""
    nm, y, xin = fx["names"], fx["y"], fx["x_physical"]
    order3, order2 = ["T", "Q", "U", "V"], ["SP", "toa_up_sw", "toa_up_lw", "sfc_dn_sw", "sfc_up_sw", "sfc_dn_lw", "sfc_up_lw",
                                             "sfc_sh", "sfc_lh", "tp", "evap"]
    packed = torch.cat([y[nm[k]] for k in order3 + order2], dim=1).to(dev).contiguous()
    L = y[nm["T"]].shape[1]
    pred3 = {k: packed[:, i * L: (i + 1) * L, 0] for i, k in enumerate(order3)}
    pred2 = {k: packed[:, 4 * L + i, 0] for i, k in enumerate(order2)}
    in_all = torch.cat([xin[nm[k]] for k in order3], dim=1).to(dev).contiguous()      # [B, 4L, T, H, W]
    in3 = {k: in_all[:, i * L: (i + 1) * L, -1] for i, k in enumerate(order3)}
    sp_in = xin[nm["SP"]].to(dev)[:, 0, -1]
    solin = xin[nm["SOLIN"]].to(dev)[:, 0, -1]
    return packed, pred3, pred2, in3, sp_in, solin


def test_water_and_energy_fixers_through_the_emulator(golden_dir, emulated, monkeypatch):
    """GlobalWaterFixerB200 / GlobalEnergyFixerB200 (host logic + documented kernel semantics) vs the UNMODIFIED reference
    classes GlobalWaterFixer / GlobalEnergyFixerUpDown (tests/golden/make_golden_fixers.py)."""
    fx = torch.load(os.path.join(golden_dir, "fixers.pt"), weights_only=False)
    _packed, pred3, pred2, in3, sp_in, solin = _fixer_views(fx, "cpu")
    hours = fx["n_seconds"] // 3600
    wf = pipeline.GlobalWaterFixerB200(fx["area"], fx["coef_a"], fx["coef_b"], hours, device="cpu")
    ratio = wf.apply(pred3["Q"], pred2["SP"], in3["Q"], sp_in, pred2["tp"], pred2["evap"])
    ref = fx["water_fixed_tp"][:, 0, 0]
    err = float((pred2["tp"] - ref).abs().max() / ref.abs().max())
    print("water fixer ratio", ratio.tolist(), "rel err", err)
    assert err < 5e-6
    # a latitude-band split of the sums adds up to the global sums (what a decomposed forecast all-reduces)
    H = ref.shape[-2]
    full = wf.sums(pred3["Q"], pred2["SP"], in3["Q"], sp_in, pred2["tp"], pred2["evap"])
    parts = (wf.sums(pred3["Q"], pred2["SP"], in3["Q"], sp_in, pred2["tp"], pred2["evap"], rows=(0, 4))
             + wf.sums(pred3["Q"], pred2["SP"], in3["Q"], sp_in, pred2["tp"], pred2["evap"], rows=(4, H - 4)))
    assert torch.allclose(full, parts, rtol=1e-12)

    ef = pipeline.GlobalEnergyFixerB200(fx["area"], fx["coef_a"], fx["coef_b"], fx["gph_surf"], hours, device="cpu")
    p2 = [pred2[k] for k in ("SP", "toa_up_sw", "toa_up_lw", "sfc_dn_sw", "sfc_up_sw", "sfc_dn_lw", "sfc_up_lw", "sfc_sh", "sfc_lh")]
    ratio = ef.apply([pred3[k] for k in ("T", "Q", "U", "V")], p2, [in3[k] for k in ("T", "Q", "U", "V")], sp_in, solin)
    ref = fx["energy_fixed_T"][:, :, 0]
    err = float((pred3["T"] - ref).abs().max() / ref.abs().max())
    print("energy fixer ratio", ratio.tolist(), "rel err", err)
    assert err < 5e-6
    assert emulated.calls.count("energy_budget_sums") == 1 and emulated.calls.count("energy_fix_temperature") == 1
