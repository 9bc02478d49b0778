"""CPU: the oracle restatement against vectors produced by the unmodified reference."""
import json
import os

import pytest
import torch

from miles_credit_b200.geometry import build_geometry, flops_per_forward, state_spec, workload
from miles_credit_b200.synth import state_checksum, synthetic_input, synthetic_state_dict
from oracle import crossformer_oracle as oracle


@pytest.mark.parametrize("case", ["unit", "unit_mirror_f2", "unit_wxformer"])
def test_oracle_matches_reference_forward(golden_dir, case):
    fx = torch.load(os.path.join(golden_dir, f"{case}.pt"), weights_only=False)
    geo = build_geometry(**fx["kwargs"])
    sd = synthetic_state_dict(geo, seed=fx["seed"])
    assert state_checksum(sd) == pytest.approx(fx["state_checksum"], rel=1e-9), "synthetic weight generator drifted"
    assert {k: list(v.shape) for k, v in sd.items()} == fx["keys"]
    x = synthetic_input(geo, batch=fx["batch"], seed=fx["seed"])
    taps = {}
    with torch.no_grad():
        y = oracle.forward(x, sd, geo, taps)
    assert y.shape == fx["y"].shape
    # fp32 reassociation only (tolerance: 1e-5 rel-max; measured 1e-6)
    assert float((y - fx["y"]).abs().max() / fx["y"].abs().max()) < 1e-5
    for name, ref in fx["taps"].items():
        assert float((taps[name] - ref).abs().max() / ref.abs().max()) < 1e-5, name


def test_padding_known_answers(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "padding.pt"), weights_only=False)
    for key, ref in fx["padded"].items():
        mode, a, b, c, d = key.split("_")
        out = oracle.pad_field(fx["x"], mode, (int(a), int(b)), (int(c), int(d)))
        assert torch.equal(out, ref), key
        assert torch.equal(oracle.unpad_field(out, (int(a), int(b)), (int(c), int(d))), fx["x"])


def test_state_spec_matches_reference_keys(golden_dir):
    keys = json.load(open(os.path.join(golden_dir, "state_keys.json")))
    for name, ref in keys.items():
        wl, _, variant = name.partition(":")
        spec = state_spec(build_geometry(**dict(workload(wl), variant=variant or "crossformer")))
        assert {k: list(s) for k, (s, _) in spec.items()} == ref


def test_geometry_of_headline_config():
    geo = build_geometry(**workload("wxformer_6h_025deg"))
    assert (geo.h_pad, geo.w_pad) == (801, 1600)
    assert [(s.h, s.w) for s in geo.stages] == [(400, 800), (200, 400), (100, 200), (50, 100)]
    assert [b.c_out for b in geo.stages[0].branches] == [64, 32, 16, 16]
    assert (geo.input_channels, geo.output_channels) == (60, 64)
    assert (geo.h_crop, geo.w_crop, geo.h_out, geo.w_out) == (720, 1440, 721, 1440)
    fl = flops_per_forward(geo)
    assert fl["total"] == pytest.approx(5.548e12, rel=2e-3)
    assert fl["ff"] == pytest.approx(2348.8e9, rel=1e-3)


def test_bilinear_matches_torch():
    x = torch.randn(2, 3, 12, 10)
    ref = torch.nn.functional.interpolate(x, size=(13, 10), mode="bilinear")
    assert torch.allclose(oracle.bilinear_resize(x, 13, 10), ref, atol=1e-6)
    ref = torch.nn.functional.interpolate(x, size=(7, 17), mode="bilinear")
    assert torch.allclose(oracle.bilinear_resize(x, 7, 17), ref, atol=1e-6)


def test_wxformer_variant_ignores_crossformer_only_keys():
    """`type: wxformer` configs of the reference carry keys its class swallows in **kwargs (wxformer/crossformer.py:624-653;
    config/example-v2026.1.0.yml:194-196, config/gen_2/smoke/smoke_gen2_multistep_casper.yml:68-69): same geometry with them."""
    kw = dict(workload("unit"), variant="wxformer")
    geo = build_geometry(**kw)
    assert build_geometry(**dict(kw, upsample_v_conv=True, decoder_attention_type="scse", frame_patch_size=2)) == geo
    with pytest.raises(NotImplementedError):  # the `crossformer` class does build a different decoder for it
        build_geometry(**dict(workload("unit"), upsample_v_conv=True))


def test_scope_errors():
    with pytest.raises(NotImplementedError):
        build_geometry(**dict(workload("unit"), patch_height=2, patch_width=2))
    with pytest.raises(NotImplementedError):
        build_geometry(**dict(workload("unit"), post_conf={"activate": True}))
    with pytest.raises(ValueError):
        build_geometry(**dict(workload("unit"), local_window_size=7))
