"""CPU: FuXi host logic (geometry, state-dict layout, weight preparation, launch plan, index lists) executed through the
C-ABI emulator and compared with the golden vectors of the UNMODIFIED reference module credit/models/fuxi.py
(tests/golden/make_golden_fuxi.py).  The CUDA kernels themselves are checked by tests/test_gpu_fuxi.py."""
import os

import pytest
import torch

from miles_credit_b200 import fuxi as wfuxi
from miles_credit_b200 import lib as wlib
from miles_credit_b200 import ops

from abi_emulator import EmulatedLib


@pytest.fixture
def emulated(monkeypatch):
    emu = EmulatedLib()
    monkeypatch.setattr(wlib, "_lib", emu)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(ops, "_req", lambda *a, **k: None)
    return emu


@pytest.mark.parametrize("case", ["unit_fuxi", "unit_fuxi_nopad"])
def test_fuxi_plan_through_emulated_abi_matches_reference(golden_dir, emulated, case):
    fx = torch.load(os.path.join(golden_dir, f"{case}.pt"), weights_only=False)
    geo = wfuxi.build_fuxi_geometry(**fx["kwargs"])
    # state-dict layout = the reference module's (fuxi.py + timm's parameter tree)
    spec = wfuxi.fuxi_state_spec(geo)
    assert {k: tuple(s) for k, (s, _) in spec.items()} == {k: tuple(v.shape) for k, v in fx["state_dict"].items()}
    wts = wfuxi.prepare_fuxi(fx["state_dict"], geo)
    plan = wfuxi._FuxiPlan(geo, wts, fx["x"].shape[0], torch.device("cpu"))
    y = plan.run(fx["x"].contiguous())
    assert y.shape == fx["y"].shape
    err = float((y - fx["y"]).abs().max() / fx["y"].abs().max())
    print(case, "rel-max vs the reference", err)
    assert err < 2e-5, err
    assert emulated.calls.count("swin_attention") == geo.depth
    assert emulated.calls.count("layernorm_residual") == 2 * geo.depth
    assert emulated.calls.count("gemm_tc") == 4 * geo.depth + 1
    assert emulated.calls.count("conv_tc") == 1 + 3 + 3          # cube, down (conv + 2), up (convT + 2)


def test_fuxi_geometry_of_the_025deg_workload():
    geo = wfuxi.build_fuxi_geometry(**wfuxi.fuxi_workload("fuxi_6h_025deg"))
    assert (geo.h_pad, geo.w_pad, geo.lat, geo.lon, geo.th, geo.tw) == (800, 1600, 200, 400, 100, 200)
    assert geo.pad2d == (1, 2, 2, 3) and (geo.gh, geo.gw) == (105, 203)   # get_pad2d: smaller half first (fuxi.py:67-79)
    assert geo.ws == (7, 7) and geo.shift == (3, 3) and geo.dh == 128
    assert (geo.in_chans, geo.out_chans) == (74, 71)
    fl = wfuxi.fuxi_flops_per_forward(geo)
    assert 13e12 < fl["total"] < 14.5e12                                   # "est. 13-14 TF / forward"
    arx = wfuxi.build_fuxi_geometry(**wfuxi.fuxi_workload("fuxi_6h_arxiv"))
    assert (arx.lat, arx.lon, arx.th, arx.tw, arx.gh, arx.gw) == (200, 360, 100, 180, 105, 182)
    with pytest.raises(NotImplementedError):
        wfuxi.build_fuxi_geometry(**dict(wfuxi.fuxi_workload("fuxi_1deg"), use_noise=True))
    with pytest.raises(ValueError):
        wfuxi.build_fuxi_geometry(**dict(wfuxi.fuxi_workload("fuxi_1deg"), image_height=180))


def test_fuxi_module_surface_and_state_dict(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "unit_fuxi.pt"), weights_only=False)
    m = wfuxi.FuxiB200(**fx["kwargs"])
    msg = m.load_state_dict(fx["state_dict"], strict=True)
    assert not msg.missing_keys and not msg.unexpected_keys and not m._lazy_init
    assert m.use_padding and m.use_interp and m.out_chans == 10 and m.patch_size == (2, 4, 4)
    assert m.input_resolution == (8, 14) and m.img_size == (2, 64, 112) and m.img_size_original == (2, 45, 88)
    with pytest.raises(RuntimeError):                                      # no CPU path
        m.eval()(fx["x"])
    m.train()
    with pytest.raises(NotImplementedError):
        m(fx["x"])
    lazy = wfuxi.FuxiB200(**fx["kwargs"])
    assert set(lazy.state_dict()) == set(fx["state_dict"]) and all(torch.isfinite(v).all() for v in lazy.state_dict().values())
