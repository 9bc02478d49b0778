"""GPU parity of the tcgen05 implicit-GEMM convolution (4-D TMA boxes) against fp64 torch convolutions."""
import pytest
import torch
import torch.nn.functional as F

from miles_credit_b200 import ops
from miles_credit_b200.weights import conv_tc_weights, conv_weights, convt_k2s2_weights, convt_k4s2p1_weights

pytestmark = pytest.mark.gpu
DEV = "cuda"


def to_pm(x):
    return x.permute(0, 2, 3, 1).contiguous()


def planes_of(x_pm):
    b, h, w, c = x_pm.shape
    hi = torch.empty(b, h, w, c, device=DEV, dtype=torch.float16)
    lo = torch.empty_like(hi)
    ops.split_f16x2(x_pm, c, hi, lo, c, b * h * w, c)
    return hi, lo


def relmax(a, b):
    return float((a - b).abs().max() / b.abs().max())


CASES = [
    # cin, cout, k, stride, pad, H, W, B
    (128, 128, 2, 2, 0, 26, 42, 1),
    (128, 64, 4, 2, 1, 26, 42, 2),
    (64, 96, 3, 1, 1, 13, 21, 2),
    (256, 256, 3, 1, 1, 10, 20, 1),
    (96, 32, 1, 1, 0, 9, 7, 1),
    (512, 512, 2, 2, 0, 20, 40, 1),
]


@pytest.mark.parametrize("case", CASES)
def test_conv_tc_matches_torch(case):
    cin, cout, k, s, p, h, w, b = case
    torch.manual_seed(cin + cout + k)
    x = torch.randn(b, cin, h, w)
    wt = torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5
    bias = torch.randn(cout)
    ref = F.conv2d(x.double(), wt.double(), bias.double(), stride=s, padding=p).float()
    ho, wo = ref.shape[-2:]
    tw = conv_tc_weights(conv_weights(wt.to(DEV), bias.to(DEV), s, p))
    hi, lo = planes_of(to_pm(x).to(DEV))
    ldc = cout + 8
    out = torch.zeros(b, ho, wo, ldc, device=DEV)
    res = torch.randn(b, ho, wo, cout, device=DEV)
    o_hi = torch.zeros(b, ho, wo, cout, device=DEV, dtype=torch.float16)
    o_lo = torch.zeros_like(o_hi)
    ops.conv_f16x2_tc(ops.make_conv_tc_desc(hi, lo, tw, B=b, Hi=h, Wi=w, lda=cin, Ho=ho, Wo=wo, out=out, ldc=ldc, c_off=8,
                                            res=res, ldr=cout, out_hi=o_hi, out_lo=o_lo, ldh=cout))
    torch.cuda.synchronize()
    want = to_pm(ref) + res.cpu()
    err = relmax(out[..., 8:].cpu(), want)
    print(f"conv_tc {case}: rel-max {err:.3e}")
    assert err < 3e-6
    assert torch.count_nonzero(out[..., :8]) == 0
    assert relmax(o_hi.float().cpu() + o_lo.float().cpu(), want) < 3e-6


@pytest.mark.parametrize("cin,cout,h,w,b", [(256, 128, 7, 11, 2), (1024, 512, 5, 10, 1)])
def test_conv_transpose_k2s2_tc(cin, cout, h, w, b):
    torch.manual_seed(cin)
    x = torch.randn(b, cin, h, w)
    wt = torch.randn(cin, cout, 2, 2) / cin**0.5
    bias = torch.randn(cout)
    ref = F.conv_transpose2d(x.double(), wt.double(), bias.double(), stride=2).float()
    tw = conv_tc_weights(convt_k2s2_weights(wt.to(DEV), bias.to(DEV)))
    hi, lo = planes_of(to_pm(x).to(DEV))
    out = torch.zeros(b, 2 * h, 2 * w, cout, device=DEV)
    ops.conv_f16x2_tc(ops.make_conv_tc_desc(hi, lo, tw, B=b, Hi=h, Wi=w, lda=cin, Ho=h, Wo=w, out=out, ldc=cout))
    err = relmax(out.cpu(), to_pm(ref))
    print(f"convT k2s2 tc {cin}->{cout}: rel-max {err:.3e}")
    assert err < 3e-6


@pytest.mark.parametrize("cin,cout,h,w,b", [(128, 64, 12, 8, 1), (256, 84, 9, 14, 2)])
def test_conv_transpose_k4s2p1_tc(cin, cout, h, w, b):
    torch.manual_seed(cout)
    x = torch.randn(b, cin, h, w)
    wt = torch.randn(cin, cout, 4, 4) / (4 * cin) ** 0.5
    bias = torch.randn(cout)
    ref = F.conv_transpose2d(x.double(), wt.double(), bias.double(), stride=2, padding=1).float()
    tw = conv_tc_weights(convt_k4s2p1_weights(wt.to(DEV), bias.to(DEV)))
    hi, lo = planes_of(to_pm(x).to(DEV))
    out = torch.zeros(b, 2 * h, 2 * w, cout, device=DEV)
    ops.conv_f16x2_tc(ops.make_conv_tc_desc(hi, lo, tw, B=b, Hi=h, Wi=w, lda=cin, Ho=h, Wo=w, out=out, ldc=cout))
    err = relmax(out.cpu(), to_pm(ref))
    print(f"convT k4s2p1 tc {cin}->{cout}: rel-max {err:.3e}")
    assert err < 3e-6


def test_conv_tc_channel_slice_input():
    """Input planes that are the upper half of a concat buffer (pointer offset + pixel stride 2C)."""
    torch.manual_seed(1)
    c, h, w = 64, 12, 10
    both = torch.randn(1, 2 * c, h, w)
    wt = torch.randn(128, c, 2, 2) / (4 * c) ** 0.5
    ref = F.conv2d(both[:, c:].double(), wt.double(), None, stride=2).float()
    hi, lo = planes_of(to_pm(both).to(DEV))
    tw = conv_tc_weights(conv_weights(wt.to(DEV), None, 2, 0))
    out = torch.zeros(1, h // 2, w // 2, 128, device=DEV)
    ops.conv_f16x2_tc(ops.make_conv_tc_desc(hi[..., c:], lo[..., c:], tw, B=1, Hi=h, Wi=w, lda=2 * c, Ho=h // 2, Wo=w // 2,
                                            out=out, ldc=128))
    assert relmax(out.cpu(), to_pm(ref)) < 3e-6


@pytest.mark.parametrize("cin,ch,k,h,w,b", [(60, 16, 32, 97, 300, 1), (60, 16, 16, 97, 300, 1), (60, 32, 8, 45, 130, 2),
                                            (60, 64, 4, 45, 130, 2), (10, 4, 32, 97, 144, 1), (24, 8, 8, 33, 64, 1)])
def test_cross_embed_toeplitz_tc(cin, ch, k, h, w, b):
    """Stage-0 cross-embed branch (stride 2, pad (k-2)/2) as the Toeplitz-lifted tensor-core GEMM vs fp64 conv2d."""
    from miles_credit_b200.weights import toeplitz_weights

    torch.manual_seed(cin * k + ch)
    x = torch.randn(b, cin, h, w)
    wt = torch.randn(ch, cin, k, k) / (cin * k * k) ** 0.5
    bias = torch.randn(ch)
    p = (k - 2) // 2
    ref = F.conv2d(x.double(), wt.double(), bias.double(), stride=2, padding=p).float()
    ho, wo = ref.shape[-2:]
    xpm = torch.zeros(b, h, w, 64)
    xpm[..., :cin] = to_pm(x)
    hi, lo = planes_of(xpm.to(DEV))
    tw = toeplitz_weights(wt.to(DEV), bias.to(DEV), p)
    ldc = 2 * ch + 8
    out = torch.zeros(b, ho, wo, ldc, device=DEV)
    ops.cross_embed_toeplitz_tc(ops.make_toeplitz_desc(hi, lo, tw, out, B=b, Hi=h, Wi=w, lda=64, Ho=ho, Wo=wo, ldc=ldc,
                                                       c_off=8))
    torch.cuda.synchronize()
    err = relmax(out[..., 8: 8 + ch].cpu(), to_pm(ref))
    print(f"toeplitz cin={cin} ch={ch} k={k}: rel-max {err:.3e}")
    assert err < 6e-6  # K = 2k*64 up to 4096 with two accumulators (see test_gpu_gemm_tc)
    assert torch.count_nonzero(out[..., :8]) == 0 and torch.count_nonzero(out[..., 8 + ch:]) == 0


def test_pad_planes_match_fp32_pad():
    torch.manual_seed(4)
    x = torch.randn(2, 5, 2, 9, 16, device=DEV)
    ref = ops.pad_to_pixel_major(x, (3, 4), (5, 2), "earth", 64)
    hi = torch.empty(2, 16, 23, 64, device=DEV, dtype=torch.float16)
    lo = torch.empty_like(hi)
    ops.pad_to_pixel_major_f16x2(x, (3, 4), (5, 2), "earth", 64, hi, lo)
    assert torch.equal(hi, ref.half())
    assert float((hi.float() + lo.float() - ref).abs().max()) < 1e-6


@pytest.mark.parametrize("cin,c,h,w,b", [(128, 64, 9, 13, 2), (256, 12, 10, 8, 1)])
def test_subpixel_conv_pixelshuffle(cin, c, h, w, b):
    """Conv2d(Cin -> 4C, 3x3) + PixelShuffle(2) as 4 output-parity phases with per-phase bias (wxformer decoder)."""
    from miles_credit_b200.weights import conv_ps_weights

    torch.manual_seed(c)
    x = torch.randn(b, cin, h, w)
    wt = torch.randn(4 * c, cin, 3, 3) / (9 * cin) ** 0.5
    bias = torch.randn(4 * c)
    ref = F.pixel_shuffle(F.conv2d(x.double(), wt.double(), bias.double(), padding=1), 2).float()
    cw = conv_ps_weights(wt.to(DEV), bias.to(DEV))
    hi, lo = planes_of(to_pm(x).to(DEV))
    out = torch.zeros(b, 2 * h, 2 * w, c, device=DEV)
    ops.conv_f16x2_tc(ops.make_conv_tc_desc(hi, lo, conv_tc_weights(cw), B=b, Hi=h, Wi=w, lda=cin, Ho=h, Wo=w, out=out, ldc=c))
    assert relmax(out.cpu(), to_pm(ref)) < 3e-6
    out32 = torch.zeros(b, 2 * h, 2 * w, c, device=DEV)
    ops.conv_igemm_f32(ops.make_conv_desc(to_pm(x).to(DEV), cw, out32, B=b, Hi=h, Wi=w, lda=cin, Ho=h, Wo=w, ldc=c))
    assert relmax(out32.cpu(), to_pm(ref)) < 3e-6
