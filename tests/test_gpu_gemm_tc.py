"""GPU parity of the tcgen05 f16x2 GEMM against fp64 on the CPU and against the exact-fp32 CUDA-core kernel."""
import pytest
import torch

from miles_credit_b200 import lib as wlib
from miles_credit_b200 import ops
from miles_credit_b200.weights import conv_weights, gemm_weights

pytestmark = pytest.mark.gpu
DEV = "cuda"


def planes(x):
    m, k = x.shape
    hi = torch.empty(m, k, device=DEV, dtype=torch.float16)
    lo = torch.empty(m, k, device=DEV, dtype=torch.float16)
    ops.split_f16x2(x, k, hi, lo, k, m, k)
    return hi, lo


def test_split_planes_carry_22_bits():
    torch.manual_seed(0)
    x = torch.randn(513, 96, device=DEV) * 3
    hi, lo = planes(x)
    rec = hi.float() + lo.float()
    assert float((rec - x).abs().max()) < 4e-7 * float(x.abs().max())
    assert torch.equal(hi, x.half())


CASES = [
    # M, N, K
    (300, 128, 128),
    (1000, 384, 128),
    (257, 512, 256),
    (128, 64, 64),
    (129, 96, 32),
    (77, 72, 200),
    (2100, 1024, 4096),
    (5000, 128, 512),
]


@pytest.mark.parametrize("m,n,k", CASES)
def test_gemm_tc_plain(m, n, k):
    torch.manual_seed(m + n + k)
    a = torch.randn(m, k)
    w = torch.randn(n, k) / k**0.5
    ref = (a.double() @ w.double().t()).float()
    a_hi, a_lo = planes(a.to(DEV))
    gw = gemm_weights(w.to(DEV), None)
    out = torch.full((m, n), float("nan"), device=DEV)
    ops.gemm_f16x2_tc(ops.make_gemm_desc(a_hi, a_lo, gw, M=m, lda=k, out=out, ldc=n))
    torch.cuda.synchronize()
    got = out.cpu()
    err = float((got - ref).abs().max() / ref.abs().max())
    print(f"gemm_tc {m}x{n}x{k}: rel-max {err:.3e}")
    assert torch.isfinite(got).all()
    # the tensor-core accumulator truncates on every accumulate; split accumulators keep K=4096 at 7e-6
    assert err < (1e-5 if k >= 2048 else 3e-6)


def test_gemm_tc_epilogues_match_exact_kernel():
    """bias + GELU -> fp16 planes, then second GEMM with bias + residual into a strided slice: the FF block."""
    torch.manual_seed(5)
    m, d = 1500, 128
    x = torch.randn(m, d)
    w1, b1 = torch.randn(4 * d, d) / d**0.5, torch.randn(4 * d) * 0.1
    w2, b2 = torch.randn(d, 4 * d) / (4 * d) ** 0.5, torch.randn(d) * 0.1
    resid = torch.randn(m, 2 * d)
    h_ref = torch.nn.functional.gelu(x.double() @ w1.double().t() + b1.double())
    y_ref = (h_ref @ w2.double().t() + b2.double()).float() + resid[:, d:]

    xd = x.to(DEV)
    a_hi, a_lo = planes(xd)
    g1, g2 = gemm_weights(w1.to(DEV), b1.to(DEV)), gemm_weights(w2.to(DEV), b2.to(DEV))
    h_hi = torch.empty(m, 4 * d, device=DEV, dtype=torch.float16)
    h_lo = torch.empty_like(h_hi)
    ops.gemm_f16x2_tc(ops.make_gemm_desc(a_hi, a_lo, g1, M=m, lda=d, out_hi=h_hi, out_lo=h_lo, ldh=4 * d, act=wlib.ACT_GELU))
    h = (h_hi.float() + h_lo.float()).cpu()
    assert float((h - h_ref.float()).abs().max() / h_ref.abs().max()) < 3e-6
    stream = resid.to(DEV).clone()
    xv = stream[:, d:]
    ops.gemm_f16x2_tc(ops.make_gemm_desc(h_hi, h_lo, g2, M=m, lda=4 * d, out=xv, ldc=2 * d, res=xv, ldr=2 * d))
    torch.cuda.synchronize()
    assert float((stream[:, d:].cpu() - y_ref).abs().max() / y_ref.abs().max()) < 3e-6
    assert torch.equal(stream[:, :d].cpu(), resid[:, :d])  # the other half of the concat buffer is untouched

    # same block through the exact-fp32 CUDA-core kernel
    c1 = conv_weights(w1.reshape(4 * d, d, 1, 1).to(DEV), b1.to(DEV), 1, 0)
    hid = torch.empty(m, 4 * d, device=DEV)
    ops.conv_igemm_f32(ops.make_conv_desc(xd, c1, hid, B=1, Hi=1, Wi=m, lda=d, Ho=1, Wo=m, ldc=4 * d, act=wlib.ACT_GELU))
    assert float((hid.cpu() - h).abs().max() / h.abs().max()) < 3e-6


@pytest.mark.parametrize("m,n,k", [(1000, 384, 128), (333, 96, 256), (5000, 512, 512), (700, 160, 2048), (129, 1536, 512)])
@pytest.mark.parametrize("bias", [False, True])
def test_gemm_tc_specialised_epilogues(m, n, k, bias):
    """The compile-time-specialised epilogues (planes, GELU -> planes, fp32, in-place residual as a TMA reduce-add store)
    of both kernel shapes (K <= 256: 16 epilogue warps; K >= 512: 8) vs fp64, ragged M and N % 128 != 0 included."""
    torch.manual_seed(m + n + k + int(bias))
    a = torch.randn(m, k)
    w = torch.randn(n, k) / k**0.5
    b = torch.randn(n) * 0.1 if bias else None
    lin = a.double() @ w.double().t() + (b.double() if bias else 0.0)
    a_hi, a_lo = planes(a.to(DEV))
    gw = gemm_weights(w.to(DEV), b.to(DEV) if bias else None)
    scale = float(lin.abs().max())
    tol = 1e-5 if k >= 2048 else 4e-6
    # fp16 planes, without and with GELU
    for act, ref in ((wlib.ACT_NONE, lin), (wlib.ACT_GELU, torch.nn.functional.gelu(lin))):
        o_hi = torch.full((m, n), float("nan"), device=DEV, dtype=torch.float16)
        o_lo = torch.full_like(o_hi, float("nan"))
        ops.gemm_f16x2_tc(ops.make_gemm_desc(a_hi, a_lo, gw, M=m, lda=k, out_hi=o_hi, out_lo=o_lo, ldh=n, act=act))
        got = (o_hi.float() + o_lo.float()).cpu().double()
        assert torch.isfinite(got).all()
        assert float((got - ref).abs().max()) < tol * scale, (act, float((got - ref).abs().max()) / scale)
    # fp32 tile into a strided slice
    wide = torch.full((m, n + 32), float("nan"), device=DEV)
    ops.gemm_f16x2_tc(ops.make_gemm_desc(a_hi, a_lo, gw, M=m, lda=k, out=wide[:, 32:], ldc=n + 32))
    assert float((wide[:, 32:].cpu().double() - lin).abs().max()) < tol * scale
    assert torch.isnan(wide[:, :32]).all()
    # in-place residual (x <- x + a w^T + b): the reduce-add store, twice on the same buffer
    res = torch.randn(m, n + 32)
    stream = res.to(DEV).clone()
    xv = stream[:, :n]
    for rep in (1, 2):
        ops.gemm_f16x2_tc(ops.make_gemm_desc(a_hi, a_lo, gw, M=m, lda=k, out=xv, ldc=n + 32, res=xv, ldr=n + 32))
        torch.cuda.synchronize()
        ref = res[:, :n].double() + rep * lin
        assert float((stream[:, :n].cpu().double() - ref).abs().max()) < rep * tol * scale + 1e-6
    assert torch.equal(stream[:, n:].cpu(), res[:, n:])


def test_gemm_tc_rejects_bad_arguments():
    a = torch.zeros(8, 12, device=DEV, dtype=torch.float16)
    gw = gemm_weights(torch.ones(8, 12, device=DEV), None)
    with pytest.raises(RuntimeError):  # K not a multiple of 8
        ops.gemm_f16x2_tc(ops.make_gemm_desc(a, a, gw, M=8, lda=12, out=torch.zeros(8, 8, device=DEV), ldc=8))
