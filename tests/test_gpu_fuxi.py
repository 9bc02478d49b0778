"""GPU parity of the FuXi path: the FuXi-specific kernels against fp64 PyTorch statements of
their documented semantics, and the whole forward against the golden vectors of the UNMODIFIED reference module
(everything FuXi owns in the reference tree; the third-party Swin-V2 stage follows oracle/swin_v2.py on both sides)."""
import math
import os

import pytest
import torch

from miles_credit_b200 import fuxi as wfuxi
from miles_credit_b200 import ops
from oracle import fuxi_oracle, swin_v2

pytestmark = pytest.mark.gpu
TOL = 1e-4  # north-star budget: rel-max vs the reference fp32 forward


def relmax(a, b):
    return float((a - b).abs().max() / b.abs().max())


def _planes(n, d):
    return torch.empty((n, d), device="cuda", dtype=torch.float16), torch.empty((n, d), device="cuda", dtype=torch.float16)


@pytest.mark.parametrize("B,H,W,heads,dh,ws,shift", [
    (1, 14, 21, 8, 128, (7, 7), (3, 3)),      # the 0.25 deg shape: 49 tokens, head dim 128, shifted
    (1, 14, 21, 8, 128, (7, 7), (0, 0)),
    (2, 9, 15, 2, 16, (3, 3), (1, 1)),        # unit_fuxi
    (2, 5, 15, 4, 8, (5, 5), (0, 2)),         # unit_fuxi_nopad: window clamped on one axis, shift on the other only
    (1, 8, 16, 3, 32, (8, 8), (4, 4)),        # 64 tokens (the kernel's limit)
])
@pytest.mark.parametrize("planes", [True, False])
def test_swin_window_attention_kernel(B, H, W, heads, dh, ws, shift, planes):
    torch.manual_seed(0)
    d, L = heads * dh, ws[0] * ws[1]
    qkv = torch.randn(B, H, W, 3 * d)
    bias = torch.randn(heads, L, L)
    scale = torch.rand(heads) * 20 + 1
    # fp64 statement of timm's shifted window attention (oracle/swin_v2.py) on precomputed bias / scale
    x = qkv.double()
    sx = torch.roll(x, shifts=(-shift[0], -shift[1]), dims=(1, 2)) if any(shift) else x
    win = swin_v2.window_partition(sx, ws).view(-1, L, 3, heads, dh).permute(2, 0, 3, 1, 4)
    q, k, v = win[0], win[1], win[2]
    att = torch.nn.functional.normalize(q, dim=-1) @ torch.nn.functional.normalize(k, dim=-1).transpose(-2, -1)
    att = att * scale.double().view(1, heads, 1, 1) + bias.double().unsqueeze(0)
    mask = swin_v2.shift_attn_mask((H, W), ws, shift)
    if mask is not None:
        nw = mask.shape[0]
        att = (att.view(-1, nw, heads, L, L) + mask.double().unsqueeze(1).unsqueeze(0)).view(-1, heads, L, L)
    o = (att.softmax(-1) @ v).transpose(1, 2).reshape(-1, ws[0], ws[1], d)
    o = swin_v2.window_reverse(o, ws, (H, W))
    ref = (torch.roll(o, shifts=shift, dims=(1, 2)) if any(shift) else o).reshape(B * H * W, d).float()

    M = B * H * W
    if planes:
        hi, lo = _planes(M, d)
        ops.swin_window_attention(qkv.cuda(), 3 * d, bias.cuda(), scale.cuda(), hi, lo, None, d, B, H, W, d, heads, ws, shift)
        got = hi.float() + lo.float()
    else:
        got = torch.empty((M, d), device="cuda")
        ops.swin_window_attention(qkv.cuda(), 3 * d, bias.cuda(), scale.cuda(), None, None, got, d, B, H, W, d, heads, ws, shift)
    err = relmax(got.cpu(), ref)
    print(f"swin attention {B}x{H}x{W} heads {heads} dh {dh} ws {ws} shift {shift}: rel-max {err:.2e}")
    assert err < 2e-6


@pytest.mark.parametrize("M,d", [(1000, 1024), (777, 32), (300, 1536), (64, 200)])
def test_layernorm_residual_kernel(M, d):
    torch.manual_seed(1)
    x, res = torch.randn(M, d) * 3 + 1, torch.randn(M, d)
    g, b = torch.randn(d), torch.randn(d)
    ref = (res.double() + torch.nn.functional.layer_norm(x.double(), (d,), g.double(), b.double(), 1e-5)).float()
    out = res.cuda().clone()
    hi, lo = _planes(M, d)
    ops.layernorm_residual(x.cuda(), d, out, d, out, d, hi, lo, d, g.cuda(), b.cuda(), M, d)   # in place on the residual
    assert relmax(out.cpu(), ref) < 2e-6
    assert relmax((hi.float() + lo.float()).cpu(), ref) < 2e-6
    out2 = torch.empty(M, d, device="cuda")
    ops.layernorm_residual(x.cuda(), d, None, 0, out2, d, None, None, 0, g.cuda(), b.cuda(), M, d)   # no residual, fp32 only
    assert relmax(out2.cpu(), (ref - res)) < 5e-6


def test_gather_rows_ex_kernel():
    torch.manual_seed(2)
    src = torch.randn(50, 64)
    idx = torch.tensor([3, -1, 49, 0, -1, 7, 7], dtype=torch.int32)
    ref = torch.where((idx >= 0)[:, None], src[idx.clamp_min(0).long()], torch.zeros(()))
    dst = torch.full((7, 64), 9.0, device="cuda")
    hi = torch.full((7, 192), 9.0, device="cuda", dtype=torch.float16)
    lo = hi.clone()
    ops.gather_rows_ex(src.cuda(), 64, idx.cuda(), dst, 64, hi, lo, 192, 64, 7, 64)
    assert torch.equal(dst.cpu(), ref)
    assert relmax((hi[:, 64:128].float() + lo[:, 64:128].float()).cpu(), ref) < 1e-6
    assert torch.all(hi[:, :64] == 9.0) and torch.all(hi[:, 128:] == 9.0)     # only the addressed column slice is written


@pytest.mark.parametrize("C,cp,interp", [(10, 10, True), (71, 71, True), (7, 8, False)])
def test_unpatchify_unpad_resize_kernel(C, cp, interp):
    torch.manual_seed(3)
    B, Lat, Lon, ph, pw = 2, 6, 10, 4, 4
    top, left, Hc, Wc = 3, 5, 17, 28
    Ho, Wo = (18, 28) if interp else (Hc, Wc)
    y = torch.randn(B, Lat, Lon, ph * pw * cp)
    img = y.view(B, Lat, Lon, ph, pw, cp)[..., :C].permute(0, 1, 3, 2, 4, 5).reshape(B, Lat * ph, Lon * pw, C).permute(0, 3, 1, 2)
    crop = img[:, :, top: top + Hc, left: left + Wc]
    ref = torch.nn.functional.interpolate(crop, size=(Ho, Wo), mode="bilinear") if interp else crop
    out = torch.empty(B, C, Ho, Wo, device="cuda")
    ops.unpatchify_unpad_resize_to_nchw(y.cuda(), out, B, C, cp, Lat, Lon, ph, pw, top, left, Hc, Wc, Ho, Wo)
    assert float((out.cpu() - ref).abs().max()) < 2e-6
    part = torch.zeros_like(out)
    ops.unpatchify_unpad_resize_to_nchw(y.cuda(), part, B, C, cp, Lat, Lon, ph, pw, top, left, Hc, Wc, Ho, Wo, rows=(4, 5))
    assert torch.equal(part[:, :, 4:9], out[:, :, 4:9]) and float(part[:, :, :4].abs().max()) == 0.0


@pytest.mark.parametrize("case", ["unit_fuxi", "unit_fuxi_nopad"])
def test_fuxi_forward_matches_reference_golden(golden_dir, case):
    fx = torch.load(os.path.join(golden_dir, f"{case}.pt"), weights_only=False)
    model = wfuxi.FuxiB200(**fx["kwargs"])
    model.load_state_dict(fx["state_dict"], strict=True)
    model = model.cuda().eval()
    y = model(fx["x"].cuda())
    assert y.shape == fx["y"].shape and y.dtype == torch.float32
    err = relmax(y.cpu(), fx["y"])
    print(f"{case}: rel-max vs the reference module = {err:.3e}")
    assert err < TOL
    assert torch.equal(y, model(fx["x"].cuda()))   # deterministic


def test_fuxi_1deg_architecture_vs_oracle():
    """The 0.25 deg architecture (dim 1024, 8 heads of 128, window 7, patch 4) at depth 4 on the 181 x 360 grid."""
    kw = wfuxi.fuxi_workload("fuxi_1deg")
    geo = wfuxi.build_fuxi_geometry(**kw)
    sd = wfuxi.synthetic_fuxi_state_dict(geo, seed=1000, sn_iters=5)
    x = wfuxi.synthetic_fuxi_input(geo, batch=1, seed=1000)
    torch.set_num_threads(os.cpu_count() or 8)
    with torch.no_grad():
        ref = fuxi_oracle.forward(x, sd, fuxi_oracle.FuxiSpec.from_kwargs(**kw))
    model = wfuxi.FuxiB200(**kw)
    model.load_state_dict(sd, strict=True)
    y = model.cuda().eval()(x.cuda())
    assert y.shape == ref.shape == (1, 71, 1, 181, 360)
    err = relmax(y.cpu(), ref)
    print(f"fuxi_1deg: rel-max vs oracle = {err:.3e}, |y| max {float(ref.abs().max()):.3f}")
    assert err < TOL


def test_history_update_kernel_and_fuxi_rollout(golden_dir):
    """frames = 2: the rollout slides the history window (trainers/rollout_utils.py:288-311) in one in-place kernel."""
    from miles_credit_b200.rollout import Rollout

    torch.manual_seed(4)
    B, C, T, H, W, n_prog, n_dyn, Cy = 2, 9, 3, 5, 7, 5, 2, 6
    x, y, f = torch.randn(B, C, T, H, W), torch.randn(B, Cy, 1, H, W), torch.randn(B, n_dyn, 1, H, W)
    ref = x.clone()
    ref[:, :, :-1] = x[:, :, 1:]
    ref[:, :n_prog, -1] = y[:, :n_prog, 0]
    ref[:, n_prog:n_prog + n_dyn, -1] = f[:, :, 0]
    xg = x.cuda()
    ops.history_update(xg, y.cuda(), f.cuda(), n_prog, n_dyn)
    assert torch.equal(xg.cpu(), ref)
    ref2 = x.clone()
    ref2[:, :, :-1] = x[:, :, 1:]
    ref2[:, :n_prog, -1] = y[:, :n_prog, 0]
    xg = x.cuda()
    ops.history_update(xg, y.cuda(), None, n_prog, 0)      # no new forcing: carried
    assert torch.equal(xg.cpu(), ref2)

    fx = torch.load(os.path.join(golden_dir, "unit_fuxi.pt"), weights_only=False)
    spec = fuxi_oracle.FuxiSpec.from_kwargs(**fx["kwargs"])
    model = wfuxi.FuxiB200(**fx["kwargs"])
    model.load_state_dict(fx["state_dict"], strict=True)
    model = model.cuda().eval()
    n_prog = spec.channels * spec.levels + spec.surface_channels
    xo, xe, xg = fx["x"].clone(), fx["x"].cuda(), fx["x"].cuda()
    eager, graph = Rollout(model), Rollout(model, graph=True)
    for _ in range(3):
        with torch.no_grad():
            yo = fuxi_oracle.forward(xo, fx["state_dict"], spec)
        nxt = xo.clone()
        nxt[:, :, :-1] = xo[:, :, 1:]
        nxt[:, :n_prog, -1] = yo[:, :n_prog, 0]
        xo = nxt
        ye = eager.step(xe)
        yg = graph.step(xg)
        assert relmax(ye.cpu(), yo) < TOL
        assert torch.equal(ye, yg)
    assert relmax(xe.cpu(), xo) < TOL and torch.equal(xe, xg)
