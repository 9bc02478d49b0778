"""GPU parity of the tcgen05 window attention (TMA gather, block-diagonal packing) against the oracle's Attention."""
import pytest
import torch

from miles_credit_b200 import ops
from miles_credit_b200.geometry import build_geometry, workload
from miles_credit_b200.synth import synthetic_state_dict
from miles_credit_b200.weights import prepare
from oracle import crossformer_oracle as oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def to_pm(x):
    return x.permute(0, 2, 3, 1).contiguous()


def attention_core_reference(qkv_pm, bias, wsz, kind, heads, dh, scale):
    """fp64 softmax(q*scale k^T + bias) v on the gathered windows; qkv_pm [B, H, W, 3d] -> [B, H, W, d]."""
    b, h, w, d3 = qkv_pm.shape
    d = d3 // 3
    nh, nw = h // wsz, w // wsz
    out = torch.zeros(b, h, w, d, dtype=torch.float64)
    q3 = qkv_pm.double()
    for bi in range(b):
        for gh in range(nh):
            for gw in range(nw):
                if kind == 0:
                    ys, xs = gh * wsz + torch.arange(wsz), gw * wsz + torch.arange(wsz)
                else:
                    ys, xs = torch.arange(wsz) * nh + gh, torch.arange(wsz) * nw + gw
                tok = q3[bi][ys][:, xs].reshape(wsz * wsz, 3, heads, dh)
                q, k, v = (tok[:, i].transpose(0, 1) for i in range(3))
                p = (q * scale @ k.transpose(1, 2) + bias.double()).softmax(-1)
                out[bi, ys[:, None], xs[None, :]] = (p @ v).transpose(0, 1).reshape(wsz, wsz, d)
    return out.float()


@pytest.mark.parametrize("wsz,kind,h,w,d,b", [(10, 0, 20, 30, 64, 1), (10, 1, 20, 30, 64, 2), (5, 1, 20, 30, 128, 1),
                                              (3, 0, 12, 18, 32, 2), (8, 1, 16, 24, 32, 1), (2, 1, 10, 14, 64, 1),
                                              (1, 1, 6, 7, 96, 1), (4, 0, 12, 20, 64, 1), (2, 1, 12, 80, 64, 2), (4, 1, 16, 40, 32, 2),
                                              (9, 0, 18, 27, 64, 1), (3, 1, 12, 48, 32, 1), (5, 1, 10, 15, 32, 1)])
def test_window_attention_tc(wsz, kind, h, w, d, b):
    torch.manual_seed(wsz * 100 + kind)
    L = wsz * wsz
    qkv = torch.randn(b, h, w, 3 * d)
    bias = torch.randn(L, L) * 0.5
    scale = 32**-0.5
    ref = attention_core_reference(qkv, bias, wsz, kind, d // 32, 32, scale)
    m = b * h * w
    q_hi = torch.empty(m, 3 * d, device=DEV, dtype=torch.float16)
    q_lo = torch.empty_like(q_hi)
    ops.split_f16x2(qkv.to(DEV), 3 * d, q_hi, q_lo, 3 * d, m, 3 * d)
    o_hi = torch.zeros(m, d, device=DEV, dtype=torch.float16)
    o_lo = torch.zeros_like(o_hi)
    tile = ops.attention_bias_tile(bias.t().contiguous().to(DEV), w, wsz, kind)
    ops.window_attention_tc(q_hi, q_lo, 3 * d, tile, o_hi, o_lo, d, b, h, w, d, 32, wsz, kind, scale)
    torch.cuda.synchronize()
    got = (o_hi.float() + o_lo.float()).cpu().reshape(b, h, w, d)
    err = float((got - ref).abs().max() / ref.abs().max())
    print(f"attention_tc wsz={wsz} kind={kind} {h}x{w} d={d}: rel-max {err:.3e}")
    assert torch.isfinite(got).all()
    assert err < 5e-6
    # and against the exact-fp32 CUDA-core kernel
    out32 = torch.zeros(m, d, device=DEV)
    ops.window_attention_f32(qkv.to(DEV), 3 * d, bias.t().contiguous().to(DEV), out32, d, b, h, w, d, 32, wsz, kind, scale)
    assert float((out32.cpu().reshape(b, h, w, d) - ref).abs().max() / ref.abs().max()) < 5e-6
