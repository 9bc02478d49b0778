"""GPU parity: every C-ABI entry point against the CPU oracle on the same seeded inputs (pytest -m gpu)."""
import os

import pytest
import torch
import torch.nn.functional as F

from miles_credit_b200 import lib as wlib
from miles_credit_b200 import ops
from miles_credit_b200.geometry import build_geometry, workload
from miles_credit_b200.synth import synthetic_input, synthetic_state_dict
from miles_credit_b200.weights import conv_weights, convt_k2s2_weights, convt_k4s2p1_weights, prepare
from oracle import crossformer_oracle as oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def relmax(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def to_pm(x):  # NCHW -> pixel-major
    return x.permute(0, 2, 3, 1).contiguous()


def to_nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("mode,lat,lon", [("earth", (3, 4), (5, 2)), ("mirror", (2, 3), (4, 4)), ("earth", (4, 4), (0, 0)),
                                          ("earth", (9, 9), (16, 16)), ("earth", (0, 0), (3, 3))])
def test_pad_bit_exact(mode, lat, lon):
    torch.manual_seed(1)
    x = torch.randn(2, 5, 2, 9, 16)
    ref = oracle.pad_field(x, mode, lat, lon)  # [B, C, T, Hp, Wp]
    ld = 12
    out = ops.pad_to_pixel_major(x.to(DEV), lat, lon, mode, ld).cpu()
    b, c, t, hp, wp = ref.shape
    ref_pm = ref.reshape(b, c * t, hp, wp).permute(0, 2, 3, 1)
    assert torch.equal(out[..., : c * t], ref_pm)
    assert torch.count_nonzero(out[..., c * t:]) == 0


@pytest.mark.parametrize("mode,lat,lon,ld,ct", [("earth", (3, 4), (5, 2), 16, (5, 2)), ("mirror", (2, 3), (4, 4), 8, (3, 2)),
                                                ("earth", (9, 9), (70, 61), 72, (35, 2)), ("earth", (4, 4), (0, 0), 64, (60, 1))])
def test_pad_vectorised_path_and_row_range(mode, lat, lon, ld, ct):
    """ld % 8 == 0 takes the 64x64-tile kernel (16-byte stores); a row range writes exactly those padded rows."""
    torch.manual_seed(4)
    c, t = ct
    x = torch.randn(2, c, t, 11, 40)
    ref = oracle.pad_field(x, mode, lat, lon)
    b, _, _, hp, wp = ref.shape
    ref_pm = ref.reshape(b, c * t, hp, wp).permute(0, 2, 3, 1)
    out = ops.pad_to_pixel_major(x.to(DEV), lat, lon, mode, ld).cpu()
    assert torch.equal(out[..., : c * t], ref_pm)
    assert torch.count_nonzero(out[..., c * t:]) == 0
    # fp16 hi/lo planes of the same pass, rows [r0, r0 + n) only
    r0, n = 2, hp - 5
    hi = torch.full((b, hp, wp, ld), 7.0, device=DEV, dtype=torch.float16)
    lo = torch.full((b, hp, wp, ld), 7.0, device=DEV, dtype=torch.float16)
    ops.pad_to_pixel_major_f16x2(x.to(DEV), lat, lon, mode, ld, hi, lo, rows=(r0, n))
    got = (hi.float() + lo.float()).cpu()
    assert torch.all(got[:, :r0] == 14.0) and torch.all(got[:, r0 + n:] == 14.0)      # untouched rows
    err = (got[:, r0: r0 + n, :, : c * t] - ref_pm[:, r0: r0 + n]).abs().max() / ref_pm.abs().max()
    assert float(err) < 2e-6                                                          # 22 significant bits
    assert torch.count_nonzero(got[:, r0: r0 + n, :, c * t:]) == 0
    part = torch.full((b, hp, wp, ld), -3.0, device=DEV)
    ops.pad_to_pixel_major(x.to(DEV), lat, lon, mode, ld, out=part, rows=(r0, n))
    assert torch.equal(part[:, r0: r0 + n].cpu(), out[:, r0: r0 + n]) and torch.all(part[:, :r0] == -3.0)


def test_pad_golden_reference_vectors(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "padding.pt"), weights_only=False)
    x = fx["x"]
    for key, ref in fx["padded"].items():
        mode, a, b_, c_, d_ = key.split("_")
        out = ops.pad_to_pixel_major(x.to(DEV), (int(a), int(b_)), (int(c_), int(d_)), mode, 6).cpu()
        bb, c, t, hp, wp = ref.shape
        assert torch.equal(out, ref.reshape(bb, c * t, hp, wp).permute(0, 2, 3, 1)), key


@pytest.mark.parametrize("d", [32, 64, 96, 128, 256, 512, 1024])
def test_layernorm(d):
    torch.manual_seed(d)
    m, ld = 777, d + 8
    x = torch.randn(m, ld) * 3 + 0.5
    g, b = torch.randn(d), torch.randn(d)
    y = torch.zeros(m, d, device=DEV)
    ops.layernorm(x.to(DEV), ld, y, d, g.to(DEV), b.to(DEV), m, d)
    ref = oracle.channel_layer_norm(x[:, :d].t().reshape(1, d, m, 1), g.reshape(1, d, 1, 1), b.reshape(1, d, 1, 1))
    ref = ref.reshape(d, m).t()
    assert torch.allclose(y.cpu(), ref, atol=2e-5, rtol=1e-5)


CONV_CASES = [
    # (Cin, Cout, k, stride, pad, H, W, B)
    (12, 16, 4, 2, 1, 21, 32, 2),
    (12, 8, 8, 2, 3, 21, 32, 1),
    (12, 4, 32, 2, 15, 37, 48, 1),
    (10, 16, 4, 2, 1, 21, 32, 1),   # Cin % 4 != 0 -> scalar loader
    (64, 96, 3, 1, 1, 12, 18, 2),
    (32, 130, 1, 1, 0, 13, 11, 2),
    (128, 256, 2, 2, 0, 12, 18, 1),
    (64, 48, 1, 1, 0, 9, 9, 1),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_igemm_f32(case):
    cin, cout, k, s, p, h, w, b = case
    torch.manual_seed(cin * 1000 + cout + k)
    x = torch.randn(b, cin, h, w)
    wt = torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5
    bias = torch.randn(cout)
    ref = F.conv2d(x.double(), wt.double(), bias.double(), stride=s, padding=p).float()
    ho, wo = ref.shape[-2:]
    cw = conv_weights(wt.to(DEV), bias.to(DEV), s, p)
    ldc = cout + 4
    out = torch.zeros(b, ho, wo, ldc, device=DEV)
    res = torch.randn(b, ho, wo, cout, device=DEV)
    d = ops.make_conv_desc(to_pm(x).to(DEV), cw, out, B=b, Hi=h, Wi=w, lda=cin, Ho=ho, Wo=wo, ldc=ldc, c_off=4, res=res,
                           ldr=cout)
    ops.conv_igemm_f32(d)
    got = out[..., 4:].cpu()
    want = to_pm(ref) + res.cpu()
    assert relmax(got, want) < 1e-5  # fp32 accumulation over K up to 12288
    assert torch.count_nonzero(out[..., :4]) == 0


def test_conv_gelu_epilogue():
    torch.manual_seed(3)
    x = torch.randn(1, 32, 10, 10)
    wt = torch.randn(128, 32, 1, 1) / 32**0.5
    bias = torch.randn(128)
    ref = F.gelu(F.conv2d(x.double(), wt.double(), bias.double())).float()
    cw = conv_weights(wt.to(DEV), bias.to(DEV), 1, 0)
    out = torch.zeros(1, 10, 10, 128, device=DEV)
    ops.conv_igemm_f32(ops.make_conv_desc(to_pm(x).to(DEV), cw, out, B=1, Hi=10, Wi=10, lda=32, Ho=10, Wo=10, ldc=128,
                                          act=wlib.ACT_GELU))
    assert relmax(out.cpu(), to_pm(ref)) < 2e-6


@pytest.mark.parametrize("cin,cout,h,w,b", [(64, 32, 6, 9, 2), (256, 128, 5, 4, 1)])
def test_conv_transpose_k2s2(cin, cout, h, w, b):
    torch.manual_seed(cin)
    x = torch.randn(b, cin, h, w)
    wt = torch.randn(cin, cout, 2, 2) / cin**0.5
    bias = torch.randn(cout)
    ref = F.conv_transpose2d(x.double(), wt.double(), bias.double(), stride=2).float()
    cw = convt_k2s2_weights(wt.to(DEV), bias.to(DEV))
    out = torch.zeros(b, 2 * h, 2 * w, cout, device=DEV)
    ops.conv_igemm_f32(ops.make_conv_desc(to_pm(x).to(DEV), cw, out, B=b, Hi=h, Wi=w, lda=cin, Ho=h, Wo=w, ldc=cout))
    assert relmax(out.cpu(), to_pm(ref)) < 2e-6


@pytest.mark.parametrize("cin,cout,h,w,b", [(64, 9, 6, 9, 2), (128, 64, 12, 8, 1)])
def test_conv_transpose_k4s2p1(cin, cout, h, w, b):
    torch.manual_seed(cout)
    x = torch.randn(b, cin, h, w)
    wt = torch.randn(cin, cout, 4, 4) / (4 * cin) ** 0.5
    bias = torch.randn(cout)
    ref = F.conv_transpose2d(x.double(), wt.double(), bias.double(), stride=2, padding=1).float()
    cw = convt_k4s2p1_weights(wt.to(DEV), bias.to(DEV))
    ldc = cout + (4 - cout % 4) % 4
    out = torch.zeros(b, 2 * h, 2 * w, ldc, device=DEV)
    ops.conv_igemm_f32(ops.make_conv_desc(to_pm(x).to(DEV), cw, out, B=b, Hi=h, Wi=w, lda=cin, Ho=h, Wo=w, ldc=ldc))
    assert relmax(out[..., :cout].cpu(), to_pm(ref)) < 2e-6


def _attention_block_reference(stage_idx, kind_idx, geo, sd, x):
    st = geo.stages[stage_idx]
    p = f"layers.{stage_idx}.1.layers.0.{kind_idx}"
    kind, wsz = ("short", st.local_window) if kind_idx == 0 else ("long", st.global_window)
    return oracle.window_attention(x, sd, p, kind, wsz, st.heads, geo.dim_head) + x


@pytest.mark.parametrize("wl,stage_idx,kind_idx", [("unit", 0, 0), ("unit", 0, 2), ("unit", 1, 2), ("unit", 2, 2),
                                                    ("unit", 3, 2), ("lws10", 0, 0), ("lws10", 0, 2), ("lws10", 1, 2)])
def test_attention_block(wl, stage_idx, kind_idx):
    """LN -> to_qkv -> window attention -> to_out + residual against the oracle's Attention restatement."""
    kw = workload("unit") if wl == "unit" else dict(
        workload("unit"), image_height=80, image_width=160, local_window_size=10, global_window_size=[10, 5, 2, 1],
        padding_conf=dict(activate=True, mode="earth", pad_lat=[40, 40], pad_lon=[80, 80]), depth=[1, 1, 1, 1])
    geo = build_geometry(**kw)
    sd = synthetic_state_dict(geo, seed=7)
    st = geo.stages[stage_idx]
    torch.manual_seed(11)
    b, d = 2, st.dim
    x = torch.randn(b, d, st.h, st.w)
    ref = _attention_block_reference(stage_idx, kind_idx, geo, sd, x)
    wts = prepare({k: v.to(DEV) for k, v in sd.items()}, geo, 12)
    att = wts.blocks[stage_idx][0][kind_idx]
    m = b * st.h * st.w
    xv = to_pm(x).to(DEV)
    ln = torch.empty(m, d, device=DEV)
    qkv = torch.empty(m, 3 * d, device=DEV)
    ops.layernorm(xv, d, ln, d, att.ln_g, att.ln_b, m, d)
    ops.conv_igemm_f32(ops.make_conv_desc(ln, att.qkv, qkv, B=b, Hi=st.h, Wi=st.w, lda=d, Ho=st.h, Wo=st.w, ldc=3 * d))
    ops.window_attention_f32(qkv, 3 * d, att.bias_t, ln, d, b, st.h, st.w, d, geo.dim_head, att.wsz, att.kind,
                             geo.dim_head**-0.5)
    ops.conv_igemm_f32(ops.make_conv_desc(ln, att.out, xv, B=b, Hi=st.h, Wi=st.w, lda=d, Ho=st.h, Wo=st.w, ldc=d, res=xv,
                                          ldr=d))
    assert relmax(xv.cpu(), to_pm(ref)) < 1e-5


def test_groupnorm_silu():
    torch.manual_seed(5)
    for (b, c, g, h, w) in [(2, 128, 32, 24, 36), (1, 512, 128, 10, 20), (1, 64, 64, 31, 17), (2, 32, 32, 48, 72)]:
        x = torch.randn(b, c, h, w) * 2 + 0.3
        gamma, beta = torch.randn(c), torch.randn(c)
        res = torch.randn(b, c, h, w)
        ref = F.silu(F.group_norm(x.double(), g, gamma.double(), beta.double(), 1e-5)).float() + res
        stats = torch.empty(b, g, 2, device=DEV)
        scratch = torch.empty(ops.groupnorm_scratch_bytes(b, h * w, c) // 4 + 4, device=DEV)
        y = torch.zeros(b, h, w, 2 * c, device=DEV)
        ops.groupnorm_silu(to_pm(x).to(DEV), c, stats, scratch, gamma.to(DEV), beta.to(DEV), to_pm(res).to(DEV), c, y,
                           2 * c, b, h * w, c, g)
        assert torch.allclose(y[..., :c].cpu(), to_pm(ref), atol=3e-5, rtol=1e-5), (b, c, g)


@pytest.mark.parametrize("hd,wd,top,left,hc,wc,ho,wo,c", [(24, 40, 3, 4, 16, 32, 17, 32, 9), (20, 20, 0, 0, 20, 20, 20, 20, 64),
                                                          (30, 50, 5, 6, 20, 36, 31, 45, 70)])
def test_unpad_resize(hd, wd, top, left, hc, wc, ho, wo, c):
    torch.manual_seed(2)
    b = 2
    y = torch.randn(b, c, hd, wd)
    ref = oracle.bilinear_resize(y[..., top: top + hc, left: left + wc], ho, wo)
    tor = F.interpolate(y[..., top: top + hc, left: left + wc], size=(ho, wo), mode="bilinear")
    assert torch.allclose(ref, tor, atol=2e-6)
    ld = c + 3
    ypm = torch.zeros(b, hd, wd, ld)
    ypm[..., :c] = to_pm(y)
    out = torch.empty(b, c, ho, wo, device=DEV)
    ops.unpad_resize_to_nchw(ypm.to(DEV), ld, out, b, c, hd, wd, top, left, hc, wc, ho, wo)
    assert torch.allclose(out.cpu(), ref, atol=2e-6)


def test_unpad_resize_row_range():
    """Only output rows [o0, o0 + n) are written (a lat-band rank writes its own rows of the prediction)."""
    torch.manual_seed(5)
    b, c, hd, wd, top, left, hc, wc, ho, wo = 1, 64, 30, 50, 5, 6, 20, 36, 21, 36
    y = torch.randn(b, c, hd, wd)
    ref = oracle.bilinear_resize(y[..., top: top + hc, left: left + wc], ho, wo)
    ypm = to_pm(y).contiguous().to(DEV)
    full = torch.empty(b, c, ho, wo, device=DEV)
    ops.unpad_resize_to_nchw(ypm, c, full, b, c, hd, wd, top, left, hc, wc, ho, wo)
    assert torch.allclose(full.cpu(), ref, atol=2e-6)
    part = torch.full((b, c, ho, wo), 9.0, device=DEV)
    ops.unpad_resize_to_nchw(ypm, c, part, b, c, hd, wd, top, left, hc, wc, ho, wo, rows=(4, 9))
    assert torch.equal(part[:, :, 4:13], full[:, :, 4:13])
    assert torch.all(part[:, :, :4] == 9.0) and torch.all(part[:, :, 13:] == 9.0)
    ops.unpad_resize_to_nchw(ypm, c, part, b, c, hd, wd, top, left, hc, wc, ho, wo, rows=(0, 0))  # empty band: no-op


def test_copy_channels():
    torch.manual_seed(9)
    x = torch.randn(2, 10, 1, 7, 9, device=DEV)
    y = torch.randn(2, 9, 1, 7, 9, device=DEV)
    frc = torch.randn(2, 2, 1, 7, 9, device=DEV)
    want = x.clone()
    want[:, :8] = y[:, :8]
    want[:, 8:10] = frc
    ops.copy_channels(x, y, [(0, 0, 8)])
    ops.copy_channels(x, frc, [(8, 0, 2)])
    assert torch.equal(x, want)


def test_bad_arguments_raise():
    with pytest.raises(RuntimeError):
        ops.window_attention_f32(torch.zeros(4, 96, device=DEV), 96, torch.zeros(1, device=DEV), torch.zeros(4, 32, device=DEV),
                                 32, 1, 2, 2, 32, 16, 1, 0, 1.0)  # dim_head != 32
    with pytest.raises(RuntimeError):
        ops.pad_to_pixel_major(torch.zeros(1, 2, 1, 4, 4), (1, 1), (1, 1), "earth", 4)  # CPU tensor
