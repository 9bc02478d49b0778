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
