"""Pins ``oracle/swin_v2.py`` (the restatement of the timm ``SwinTransformerV2Stage`` FuXi instantiates) against an
independent public implementation of the same algorithm: HuggingFace ``Swinv2Stage`` (golden vectors made by
``tests/golden/make_golden_swin_hf.py``; the live comparison runs wherever ``transformers`` is importable)."""
import importlib.util
import os
import sys

import pytest
import torch

from oracle import fuxi_oracle, swin_v2

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-5  # fp32 on both sides, different association of the same arithmetic


def _cases():
    return torch.load(os.path.join(GOLDEN, "swin_v2_hf.pt"), weights_only=False)["cases"]


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("name", ["w7_14x21", "w4_8x12", "w7_14x14_dh32"])
def test_oracle_stage_functions_match_hf_golden(name):
    c = _cases()[name]
    blocks = fuxi_oracle.swin_blocks({"s." + k: v for k, v in c["state_dict"].items()}, "s", c["depth"])
    with torch.no_grad():
        y = swin_v2.stage_forward(c["x"], blocks, c["heads"], tuple(c["resolution"]), c["window"])
    assert _rel(y, c["y"]) < TOL


@pytest.mark.parametrize("name", ["w7_14x21", "w4_8x12", "w7_14x14_dh32"])
def test_stub_module_matches_hf_golden(name):
    """The module-form twin that stands in for timm when the unmodified fuxi.py is imported (golden generation)."""
    c = _cases()[name]
    stage = swin_v2.SwinTransformerV2StageStub(c["dim"], c["dim"], tuple(c["resolution"]), c["depth"], c["heads"], c["window"]).eval()
    missing = stage.load_state_dict(c["state_dict"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    with torch.no_grad():
        y = stage(c["x"])
    assert _rel(y, c["y"]) < TOL


@pytest.mark.skipif(importlib.util.find_spec("transformers") is None, reason="transformers not installed")
def test_live_hf_stage_matches_oracle():
    sys.path.insert(0, GOLDEN)
    try:
        import make_golden_swin_hf as gen
    finally:
        sys.path.pop(0)
    dim, heads, res, window, depth = 40, 5, (10, 15), 5, 3
    stage = gen.hf_stage(dim, heads, res, window, depth, seed=7)
    torch.manual_seed(8)
    x = torch.randn(2, res[0], res[1], dim)
    with torch.no_grad():
        ref = stage(x.reshape(2, -1, dim), res)[0].reshape(2, res[0], res[1], dim)
        sd = gen.hf_to_timm(stage.state_dict(), depth)
        blocks = fuxi_oracle.swin_blocks({"s." + k: v for k, v in sd.items()}, "s", depth)
        y = swin_v2.stage_forward(x, blocks, heads, res, window)
    assert _rel(y, ref) < TOL
