"""CPU: the drop-in boundary (module surface, state-dict compatibility, C-ABI exports, loud failures)."""
import ctypes
import json
import os

import pytest
import torch

from miles_credit_b200 import lib as wlib
from miles_credit_b200.geometry import build_geometry, workload
from miles_credit_b200.model import CrossFormerB200
from miles_credit_b200.synth import synthetic_state_dict


def test_state_dict_is_reference_compatible(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "unit.pt"), weights_only=False)
    model = CrossFormerB200(**fx["kwargs"])
    assert {k: list(v.shape) for k, v in model.state_dict().items()} == fx["keys"]
    sd = synthetic_state_dict(build_geometry(**fx["kwargs"]), seed=fx["seed"])
    msg = model.load_state_dict(sd, strict=True)
    assert not msg.missing_keys and not msg.unexpected_keys
    assert torch.equal(model.state_dict()["layers.0.0.convs.3.weight_orig"], sd["layers.0.0.convs.3.weight_orig"])
    with pytest.raises(RuntimeError):
        model.load_state_dict(dict(sd, bogus=torch.zeros(1)), strict=True)


def test_wxformer_variant_state_dict(golden_dir):
    from miles_credit_b200.model import WXFormerB200

    fx = torch.load(os.path.join(golden_dir, "unit_wxformer.pt"), weights_only=False)
    kw = {k: v for k, v in fx["kwargs"].items() if k != "variant"}
    model = WXFormerB200(**kw, upsample_with_ps=True)
    assert {k: list(v.shape) for k, v in model.state_dict().items()} == fx["keys"]
    assert "up_block1.sharp.weight_orig" in fx["keys"] and "layers.0.0.convs.0.1.weight_orig" in fx["keys"]


def test_module_surface():
    kw = workload("unit")
    m = CrossFormerB200(**kw)
    assert m.use_padding and m.padding_opt.pad_NS == [25, 27] and m.use_interp
    assert (m.image_height, m.image_width, m.channels, m.levels, m.surface_channels) == (45, 96, 2, 3, 2)
    assert (m.input_channels, m.output_channels) == (10, 9)
    t1, t2 = m.split_and_reshape(torch.zeros(1, 9, 1, 45, 96))
    assert t1.shape == (1, 2, 3, 1, 45, 96) and t2.shape == (1, 2, 1, 45, 96)


def test_no_cpu_fallback():
    m = CrossFormerB200(**workload("unit")).eval()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 10, 1, 45, 96))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 11, 1, 45, 96))
    m.train()
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 10, 1, 45, 96))


def test_library_exports_every_declared_symbol():
    if not os.path.isfile(wlib.LIB_PATH):
        import __graft_entry__ as entry

        entry.build()
    names = wlib.declared_symbols()
    assert "wxf_conv_igemm_f32" in names and len(names) >= 10
    raw = ctypes.CDLL(wlib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/wxformer_b200.h but not exported"
    L = wlib.load()
    assert L.wxf_abi_version() == wlib.WXF_ABI_VERSION
    assert set(wlib._SIGNATURES) <= set(names)


def test_credit_registry_plugin(tmp_path):
    """With CREDIT present (build container): `type: crossformer_b200` + custom_models goes through load_model()."""
    import subprocess
    import sys

    if not os.path.isdir("/root/reference/credit"):
        pytest.skip("CREDIT sources not mounted (GPU box)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = f"""
import sys, types, torch
from torch import nn
sys.path.insert(0, "/root/reference"); sys.path.insert(0, {root!r})
stub = types.ModuleType("credit.postblock.gen1")
class PostBlock(nn.Module): pass
stub.PostBlock = PostBlock
sys.modules["credit.postblock.gen1"] = stub
from credit.models import load_model
from credit.models.base_model import BaseModel
from miles_credit_b200.geometry import workload
conf = {{"custom_models": [{os.path.join(root, "miles_credit_b200", "credit_plugin.py")!r}],
        "model": dict(workload("unit"), type="crossformer_b200")}}
m = load_model(conf)
assert isinstance(m, BaseModel) and type(m).__name__ == "CrossFormerB200", type(m)
m2 = load_model({{"custom_models": conf["custom_models"], "model": dict(workload("unit"), type="wxformer_b200")}})
ref2 = load_model({{"model": dict(workload("unit"), type="wxformer")}})
assert {{k: tuple(v.shape) for k, v in ref2.state_dict().items()}} == {{k: tuple(v.shape) for k, v in m2.state_dict().items()}}
ref = load_model({{"model": dict(workload("unit"), type="crossformer")}})
assert {{k: tuple(v.shape) for k, v in ref.state_dict().items()}} == {{k: tuple(v.shape) for k, v in m.state_dict().items()}}
m.load_state_dict(ref.state_dict(), strict=True)
# the other classes of the path registered by the same plugin file
from miles_credit_b200.fuxi import fuxi_workload
fk = dict(fuxi_workload("fuxi_1deg"), dim=32, num_groups=4, num_heads=2, depth=2)
mf = load_model({{"custom_models": conf["custom_models"], "model": dict(fk, type="fuxi_b200")}})
assert isinstance(mf, BaseModel) and type(mf).__name__ == "FuxiB200"
ek = dict(workload("unit"), noise_latent_dim=8)
me = load_model({{"custom_models": conf["custom_models"], "model": dict(ek, type="crossformer-ensemble_b200")}})
re_ = load_model({{"model": dict(ek, type="crossformer-ensemble")}})
assert {{k: tuple(v.shape) for k, v in re_.state_dict().items()}} == {{k: tuple(v.shape) for k, v in me.state_dict().items()}}
print("ok")
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_lazy_initialisation_and_checkpoint_load(tmp_path):
    """Constructing the module costs no weight synthesis: the synthetic initial weights appear at the first state_dict() /
    forward unless a checkpoint arrives first (what BaseModel.load_model does, base_model.py:72-85)."""
    kw = workload("unit")
    torch.manual_seed(7)
    a = CrossFormerB200(**kw)
    assert a._lazy_init
    sd_a = a.state_dict()          # materialises
    assert not a._lazy_init and all(torch.isfinite(v).all() for v in sd_a.values())
    torch.manual_seed(7)
    b = CrossFormerB200(**kw, init_weights=True)   # eager: same weights for the same torch seed
    assert not b._lazy_init and all(torch.equal(sd_a[k], v) for k, v in b.state_dict().items())
    c = CrossFormerB200(**kw)
    sd = synthetic_state_dict(build_geometry(**kw), seed=3)
    c.load_state_dict(sd, strict=True)
    assert not c._lazy_init and torch.equal(c.state_dict()["up_block4.weight_orig"], sd["up_block4.weight_orig"])
    # partial checkpoint: the missing tensors still get their initial values at first use
    d = CrossFormerB200(**kw)
    part = {k: v for k, v in sd.items() if not k.startswith("up_block4")}
    msg = d.load_state_dict(part, strict=False)
    assert msg.missing_keys and d._lazy_init
    # save_model / load_model round trip through the BaseModel classmethods
    conf = {"save_loc": str(tmp_path), "model": dict(kw, type="crossformer_b200"), "trainer": {"mode": "none"}}
    c.save_model(conf)
    e = CrossFormerB200.load_model(conf)
    assert torch.equal(e.state_dict()["layers.2.1.layers.1.0.to_qkv.weight_orig"], sd["layers.2.1.layers.1.0.to_qkv.weight_orig"])
    f = CrossFormerB200.load_model_name(conf, "checkpoint.pt")
    assert torch.equal(f.state_dict()["up_block1.b.1.weight"], sd["up_block1.b.1.weight"])
    with pytest.raises(ValueError):
        CrossFormerB200.load_model_name(conf, "nope.pt")
    bad = dict(sd, bogus=torch.zeros(1))
    torch.save({"model_state_dict": bad}, os.path.join(str(tmp_path), "checkpoint.pt"))
    with pytest.raises(RuntimeError):  # models/checkpoint.py:25-31: unexpected keys raise
        CrossFormerB200.load_model(conf)


def test_wxformer_legacy_checkpoint_keys_are_migrated():
    """Reference behaviour (wxformer/crossformer.py:239-310, tests/test_legacy_checkpoint_compat.py:62-106): cross-embed
    conv keys of pre-ZeroPad2d checkpoints are renamed on load; a ConvTranspose2d-decoder checkpoint is refused."""
    from miles_credit_b200.model import WXFormerB200

    kw = dict(workload("unit"), depth=[1, 1, 1, 1])
    geo = build_geometry(**dict(kw, variant="wxformer"))
    sd = synthetic_state_dict(geo, seed=5)
    legacy = {}
    for k, v in sd.items():
        parts = k.split(".")
        if len(parts) > 5 and parts[0] == "layers" and parts[2] == "0" and parts[3] == "convs" and parts[5] == "1":
            k = ".".join(parts[:5] + parts[6:])
        legacy[k] = v
    assert "layers.0.0.convs.2.weight_orig" in legacy and "layers.0.0.convs.2.1.weight_orig" not in legacy
    m = WXFormerB200(**kw)
    msg = m.load_state_dict(legacy, strict=True)
    assert not msg.missing_keys and not msg.unexpected_keys
    assert torch.equal(m.state_dict()["layers.0.0.convs.2.1.weight_orig"], sd["layers.0.0.convs.2.1.weight_orig"])
    m.load_state_dict(sd, strict=True)  # idempotent on the current layout
    with pytest.raises(RuntimeError):
        m.load_state_dict(dict(sd, **{"up_block4.weight": torch.zeros(1)}), strict=False)


def test_kernel_limits_are_checked_at_construction():
    kw = workload("unit")
    with pytest.raises(NotImplementedError):
        CrossFormerB200(**dict(kw, dim_head=16))
    with pytest.raises(NotImplementedError):  # 12 x 12 = 144 tokens per local window
        CrossFormerB200(**dict(kw, local_window_size=12, image_height=192, image_width=192,
                               padding_conf=dict(activate=False)))


def test_stale_library_is_refused(monkeypatch, tmp_path):
    if not os.path.isfile(wlib.LIB_PATH):
        pytest.skip("library not built")
    monkeypatch.setattr(wlib, "_lib", None)
    monkeypatch.setattr(wlib, "WXF_ABI_VERSION", wlib.WXF_ABI_VERSION + 1)
    with pytest.raises(RuntimeError, match="ABI version"):
        wlib.load()


def test_domain_layout_rejects_wide_stage_halos():
    from miles_credit_b200.domain import DomainLayout

    kw = dict(workload("unit"), cross_embed_kernel_sizes=[[4, 8, 16, 32], [2, 4, 8], [2, 4], [2, 4]])
    with pytest.raises(NotImplementedError, match="halo"):
        DomainLayout(build_geometry(**kw), 2)
