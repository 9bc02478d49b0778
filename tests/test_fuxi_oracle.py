"""CPU: the FuXi oracle (oracle/fuxi_oracle.py) against the golden vectors produced by the UNMODIFIED reference module
``credit/models/fuxi.py`` (tests/golden/make_golden_fuxi.py).  The in-tree parts of FuXi are pinned by these vectors; the
Swin-V2 stage is third-party ``timm`` code that is absent here, so BOTH sides use the restatement of oracle/swin_v2.py
(pinned against HuggingFace's independent Swinv2Stage in tests/test_swin_v2_vs_hf.py; timm itself is in no image here) and the
checks below on it are self-consistency and invariants."""
import os

import pytest
import torch

from oracle import fuxi_oracle, swin_v2


@pytest.mark.parametrize("case", ["unit_fuxi", "unit_fuxi_nopad"])
def test_fuxi_oracle_matches_reference_golden(golden_dir, case):
    fx = torch.load(os.path.join(golden_dir, f"{case}.pt"), weights_only=False)
    spec = fuxi_oracle.FuxiSpec.from_kwargs(**fx["kwargs"])
    with torch.no_grad():
        y = fuxi_oracle.forward(fx["x"], fx["state_dict"], spec)
    assert y.shape == fx["y"].shape == (2, spec.out_chans, 1, spec.image_height, spec.image_width)
    err = float((y - fx["y"]).abs().max() / fx["y"].abs().max())
    print(case, "oracle vs reference rel-max", err)
    assert err < 1e-6
    # the reference's state-dict layout the future drop-in module has to keep (fuxi.py + timm's parameter tree)
    keys = set(fx["state_dict"])
    for k in ("cube_embedding.proj.weight", "cube_embedding.norm.weight", "u_transformer.down.conv.weight_orig",
              "u_transformer.down.b.1.weight", "u_transformer.layer.blocks.0.attn.qkv.weight_orig",
              "u_transformer.layer.blocks.0.attn.logit_scale", "u_transformer.layer.blocks.0.attn.cpb_mlp.2.weight_u",
              "u_transformer.layer.blocks.1.mlp.fc2.weight_v", "u_transformer.up.conv.weight_orig", "fc.weight_orig",
              "fc.bias"):
        assert k in keys, k
    assert "cube_embedding.proj.weight_orig" not in keys  # the spectral-norm hook skips Conv3d (fuxi.py:17-23)


def test_fuxi_qkv_uses_unnormalised_weight(golden_dir):
    """Reference quirk kept by the oracle: timm never calls the qkv module, so its spectral-norm hook never fires and the
    projection runs on weight_orig.  Normalising it (what a naive fold would do) changes the output."""
    fx = torch.load(os.path.join(golden_dir, "unit_fuxi.pt"), weights_only=False)
    spec = fuxi_oracle.FuxiSpec.from_kwargs(**fx["kwargs"])
    sd = dict(fx["state_dict"])
    from oracle.crossformer_oracle import effective_weight

    for i in range(spec.depth):
        p = f"u_transformer.layer.blocks.{i}.attn.qkv"
        sd[p + ".weight_orig"] = effective_weight(fx["state_dict"], p)  # as if the hook had fired
    with torch.no_grad():
        y = fuxi_oracle.forward(fx["x"], sd, spec)
    assert float((y - fx["y"]).abs().max() / fx["y"].abs().max()) > 1e-3


def test_pad2d_matches_reference_rule():
    # fuxi.py:25-79: smaller half first (top / left), remainder at the bottom / right
    assert fuxi_oracle.get_pad2d((100, 180), (7, 7)) == (1, 1, 2, 3)   # 100 -> 105, 180 -> 182
    assert fuxi_oracle.get_pad2d((8, 14), (3, 3)) == (0, 1, 0, 1)
    assert fuxi_oracle.get_pad2d((14, 21), (7, 7)) == (0, 0, 0, 0)


@pytest.mark.parametrize("ws", [(3, 3), (7, 7), (4, 6)])
def test_swin_index_tables(ws):
    n = ws[0] * ws[1]
    idx = swin_v2.relative_position_index(ws)
    assert idx.shape == (n, n) and int(idx.min()) == 0 and int(idx.max()) == (2 * ws[0] - 1) * (2 * ws[1] - 1) - 1
    centre = (ws[0] - 1) * (2 * ws[1] - 1) + ws[1] - 1
    assert torch.all(idx.diagonal() == centre)            # zero offset
    assert torch.all(idx + idx.t() == 2 * centre)         # (dy, dx) <-> (-dy, -dx)
    tab = swin_v2.relative_coords_table(ws)
    assert tab.shape == (1, 2 * ws[0] - 1, 2 * ws[1] - 1, 2)
    assert float(tab.abs().max()) == pytest.approx(1.0566, abs=1e-3)  # log2(9) / log2(8) at the window edge
    assert torch.allclose(tab, -tab.flip(1, 2))           # odd in both offsets


def test_swin_shift_mask_and_windows():
    res, ws, ss = (9, 15), (3, 3), (1, 1)
    m = swin_v2.shift_attn_mask(res, ws, ss)
    nw = (res[0] // ws[0]) * (res[1] // ws[1])
    assert m.shape == (nw, 9, 9) and set(m.unique().tolist()) == {-100.0, 0.0}
    assert torch.all(m[0] == 0)                          # interior window: nothing masked
    assert torch.any(m[-1] != 0) and torch.all(m[-1].diagonal() == 0) and torch.equal(m[-1], m[-1].t())
    assert swin_v2.shift_attn_mask(res, ws, (0, 0)) is None
    x = torch.randn(2, 9, 15, 4)
    assert torch.equal(swin_v2.window_reverse(swin_v2.window_partition(x, ws), ws, res), x)
    # a window larger than the grid is clamped and its shift dropped on that axis
    assert swin_v2.clamp_window((4, 12), (5, 5), (2, 2)) == ((4, 5), (0, 2))


def test_swin_functional_and_module_forms_agree():
    torch.manual_seed(3)
    stage = swin_v2.SwinTransformerV2StageStub(16, 16, (6, 9), 3, 2, 3).eval()
    for p in stage.parameters():
        torch.nn.init.normal_(p, std=0.3)
    x = torch.randn(2, 6, 9, 16)
    blocks = []
    for b in stage.blocks:
        a = b.attn
        blocks.append(dict(qkv_w=a.qkv.weight, q_bias=a.q_bias, v_bias=a.v_bias, logit_scale=a.logit_scale,
                           cpb0_w=a.cpb_mlp[0].weight, cpb0_b=a.cpb_mlp[0].bias, cpb2_w=a.cpb_mlp[2].weight,
                           proj_w=a.proj.weight, proj_b=a.proj.bias, norm1_w=b.norm1.weight, norm1_b=b.norm1.bias,
                           fc1_w=b.mlp.fc1.weight, fc1_b=b.mlp.fc1.bias, fc2_w=b.mlp.fc2.weight, fc2_b=b.mlp.fc2.bias,
                           norm2_w=b.norm2.weight, norm2_b=b.norm2.bias))
    with torch.no_grad():
        y_mod = stage(x)
        y_fun = swin_v2.stage_forward(x, blocks, 2, (6, 9), 3)
    assert torch.allclose(y_mod, y_fun, atol=1e-6)
    # a shifted block really mixes across the seam: moving one token changes windows on both sides of the wrap
    x2 = x.clone()
    x2[:, 0, 0] += 1.0
    with torch.no_grad():
        d = (swin_v2.stage_forward(x2, blocks, 2, (6, 9), 3) - y_fun).abs().amax(dim=(0, 3))
    assert float(d[0, 0]) > 0 and int((d > 1e-7).sum()) > 9
