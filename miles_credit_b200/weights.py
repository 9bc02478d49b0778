"""Load-time weight preparation: everything input-independent is folded once, not per forward.

* spectral norm: ``W = weight_orig / (u . W_mat v)`` (eval-mode torch.nn.utils.spectral_norm; the reference
  recomputes it on every forward for 84 modules, crossformer.py:23-26, 576-578)
* dynamic position bias: the [L, L] table of every Attention (crossformer.py:158-176, 238-245, 279-286),
  including the reference's (2w-1)-stride index quirk
* weight re-layout for the implicit-GEMM kernels: [N, taps*Cin] with the channel index fastest, transposed-conv
  weights split into 4 output-parity phases, plus the (dy, dx) tap tables.

This runs in PyTorch on whatever device the parameters live on; it is not part of the per-step path.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .geometry import Geometry


def fold_spectral_norm(sd: Dict[str, torch.Tensor], prefix: str, sn_dim: int = 0) -> torch.Tensor:
    if prefix + ".weight_orig" not in sd:
        return sd[prefix + ".weight"].float()
    w = sd[prefix + ".weight_orig"].float()
    wm = w if sn_dim == 0 else w.transpose(0, sn_dim)
    wm = wm.reshape(wm.shape[0], -1)
    sigma = torch.dot(sd[prefix + ".weight_u"].float(), torch.mv(wm, sd[prefix + ".weight_v"].float()))
    return w / sigma


def position_bias_table(sd, prefix: str, wsz: int) -> torch.Tensor:
    """bias[i, j] for tokens i, j of a wsz x wsz window (fp32, [L, L])."""
    dev = sd[prefix + ".dpb.layers.0.bias"].device
    pos = torch.arange(-wsz, wsz + 1, device=dev, dtype=torch.float32)
    t = torch.stack(torch.meshgrid(pos, pos, indexing="ij"), dim=-1).reshape(-1, 2)
    for lin, ln in ((0, 1), (3, 4), (6, 7)):
        t = F.linear(t, fold_spectral_norm(sd, f"{prefix}.dpb.layers.{lin}"), sd[f"{prefix}.dpb.layers.{lin}.bias"].float())
        t = F.layer_norm(t, (t.shape[-1],), sd[f"{prefix}.dpb.layers.{ln}.weight"].float(),
                         sd[f"{prefix}.dpb.layers.{ln}.bias"].float(), 1e-5)
        t = torch.relu(t)
    t = F.linear(t, fold_spectral_norm(sd, f"{prefix}.dpb.layers.9"), sd[f"{prefix}.dpb.layers.9.bias"].float())
    t = t.reshape(-1)
    p = torch.arange(wsz, device=dev)
    tok = torch.stack(torch.meshgrid(p, p, indexing="ij"), dim=-1).reshape(-1, 2)
    rel = tok[:, None, :] - tok[None, :, :] + (wsz - 1)
    idx = rel[..., 0] * (2 * wsz - 1) + rel[..., 1]  # reference stride quirk kept
    return t[idx]


@dataclass
class ConvWeights:
    """One implicit-GEMM contraction: weights [phases, N, T*Cin], taps [phases, T, 2], bias [N] or None."""

    w: torch.Tensor
    taps: torch.Tensor
    bias: Optional[torch.Tensor]
    n: int
    t: int
    cin: int
    stride: int = 1
    phases: int = 1
    out_scale: int = 1
    bias_phase_stride: int = 0


def _pad_cin(w_ntc: torch.Tensor, cin_pad: int) -> torch.Tensor:
    """[N, T, Cin] -> [N, T, cin_pad] zero padded."""
    n, t, c = w_ntc.shape
    if cin_pad == c:
        return w_ntc
    out = w_ntc.new_zeros((n, t, cin_pad))
    out[:, :, :c] = w_ntc
    return out


def conv_weights(w: torch.Tensor, bias, stride: int, pad: int, cin_pad: Optional[int] = None) -> ConvWeights:
    """Conv2d weight [N, Cin, k, k] -> implicit-GEMM layout; taps carry (ky - pad, kx - pad)."""
    n, c, kh, kw = w.shape
    cin_pad = cin_pad or c
    w_ntc = w.permute(0, 2, 3, 1).reshape(n, kh * kw, c)
    w_ntc = _pad_cin(w_ntc, cin_pad)
    ky, kx = torch.meshgrid(torch.arange(kh), torch.arange(kw), indexing="ij")
    taps = torch.stack([ky.reshape(-1) - pad, kx.reshape(-1) - pad], dim=-1).to(torch.int32)
    return ConvWeights(w_ntc.reshape(1, n, kh * kw * cin_pad).contiguous(), taps.reshape(1, kh * kw, 2).contiguous().to(w.device),
                       None if bias is None else bias.float().contiguous(), n, kh * kw, cin_pad, stride)


def convt_k2s2_weights(w: torch.Tensor, bias) -> ConvWeights:
    """ConvTranspose2d(k=2, s=2) weight [Cin, Cout, 2, 2]: phase (dy, dx) is a 1x1 conv to output (2y+dy, 2x+dx)."""
    cin, cout = w.shape[:2]
    wz = w.permute(2, 3, 1, 0).reshape(4, cout, cin).contiguous()  # [(dy,dx), co, ci]
    taps = torch.zeros((4, 1, 2), dtype=torch.int32, device=w.device)
    return ConvWeights(wz, taps, bias.float().contiguous(), cout, 1, cin, 1, 4, 2)


def convt_k4s2p1_weights(w: torch.Tensor, bias) -> ConvWeights:
    """ConvTranspose2d(k=4, s=2, p=1) weight [Cin, Cout, 4, 4] as 4 output-parity 2x2 convolutions.

    oy = 2*iy - 1 + ky.  Even oy = 2y: (ky, iy) in {(1, y), (3, y-1)}; odd oy = 2y+1: {(0, y+1), (2, y)}.
    """
    cin, cout = w.shape[:2]
    sel = {0: ((1, 0), (3, -1)), 1: ((0, 1), (2, 0))}  # parity -> ((k, d), (k, d))
    wz = w.new_zeros((4, cout, 4, cin))
    taps = torch.zeros((4, 4, 2), dtype=torch.int32)
    for py in (0, 1):
        for px in (0, 1):
            z = py * 2 + px
            for ty, (ky, dy) in enumerate(sel[py]):
                for tx, (kx, dx) in enumerate(sel[px]):
                    t = ty * 2 + tx
                    wz[z, :, t, :] = w[:, :, ky, kx].t()
                    taps[z, t, 0], taps[z, t, 1] = dy, dx
    return ConvWeights(wz.reshape(4, cout, 4 * cin).contiguous(), taps.to(w.device), bias.float().contiguous(), cout, 4, cin, 1, 4, 2)


@dataclass
class GemmWeights:
    """Tensor-core operand planes of a 1x1-conv weight: fp16 hi/lo of W * 2^scale_log2, [N, K] row-major."""

    w_hi: torch.Tensor
    w_lo: torch.Tensor
    bias: Optional[torch.Tensor]
    n: int
    k: int
    scale_log2: int


def gemm_weights(w: torch.Tensor, bias) -> GemmWeights:
    """Split a [N, K(,1,1)] fp32 weight into fp16 hi/lo planes after a power-of-two pre-scale.

    The scale puts max|W| in [512, 1024) so the lo plane stays in fp16's normal range for every element within
    2^-13 of the largest; the GEMM epilogue multiplies by the exact inverse.
    """
    w2 = w.reshape(w.shape[0], -1).float()
    amax = float(w2.abs().max())
    k = 0 if amax == 0.0 else int(math.floor(math.log2(1024.0 / amax)))
    k = max(min(k, 24), -24)
    ws = w2 * (2.0 ** k)
    hi = ws.half()
    lo = (ws - hi.float()).half()
    return GemmWeights(hi.contiguous(), lo.contiguous(), None if bias is None else bias.float().contiguous(),
                       w2.shape[0], w2.shape[1], k)


@dataclass
class ConvTcWeights:
    """Tensor-core operand planes of a convolution: [phases, N, T*cin_pad] fp16 hi/lo of W * 2^scale_log2."""

    w_hi: torch.Tensor
    w_lo: torch.Tensor
    taps_host: object  # ctypes int32 array [phases*T*2], read by the host at launch
    bias: Optional[torch.Tensor]
    n: int
    t: int
    cin: int
    cin_pad: int
    stride: int
    phases: int
    out_scale: int
    scale_log2: int
    bias_phase_stride: int = 0


def conv_tc_weights(cw: "ConvWeights") -> ConvTcWeights:
    """Re-lay a ConvWeights ([phases, N, T*Cin] fp32) as channel-padded, pre-scaled fp16 hi/lo planes."""
    import ctypes

    cin_pad = (cw.cin + 63) // 64 * 64
    w = cw.w.reshape(cw.phases, cw.n, cw.t, cw.cin).float()
    if cin_pad != cw.cin:
        wp = w.new_zeros((cw.phases, cw.n, cw.t, cin_pad))
        wp[..., : cw.cin] = w
        w = wp
    amax = float(w.abs().max())
    k = 0 if amax == 0.0 else int(math.floor(math.log2(1024.0 / amax)))
    k = max(min(k, 24), -24)
    ws = (w * (2.0 ** k)).reshape(cw.phases, cw.n, cw.t * cin_pad)
    hi = ws.half()
    lo = (ws - hi.float()).half()
    flat = [int(v) for v in cw.taps.reshape(-1).cpu().tolist()]
    taps = (ctypes.c_int32 * len(flat))(*flat)
    return ConvTcWeights(hi.contiguous(), lo.contiguous(), taps, cw.bias, cw.n, cw.t, cw.cin, cin_pad, cw.stride, cw.phases,
                         cw.out_scale, k, cw.bias_phase_stride)


@dataclass
class ToeplitzWeights:
    """Stage-0 cross-embed branch for wxf_cross_embed_toeplitz_tc: rows (j, c), columns (ky, r, ci<64)."""

    w_hi: torch.Tensor
    w_lo: torch.Tensor
    bias: Optional[torch.Tensor]
    ch: int
    cin: int
    kernel: int
    pad: int
    scale_log2: int


def toeplitz_eligible(c_out: int, c_in: int, kernel: int, stride: int) -> bool:
    return (stride == 2 and kernel % 2 == 0 and 2 * kernel <= 64 and (kernel // 2) * c_out <= 256 and c_out % 4 == 0
            and c_in <= 64)


def toeplitz_weights(w: torch.Tensor, bias, pad: int) -> ToeplitzWeights:
    """Conv2d weight [ch, Cin, k, k] (stride 2) -> Wt[j*ch + c, (ky*2 + r)*64 + ci] = w[c, ci, ky, 2j + r]."""
    ch, cin, k, _ = w.shape
    J = k // 2
    w5 = w.float().reshape(ch, cin, k, J, 2).permute(3, 0, 2, 4, 1)  # j, c, ky, r, ci
    wp = w5.new_zeros((J, ch, k, 2, 64))
    wp[..., :cin] = w5
    amax = float(wp.abs().max())
    kk = 0 if amax == 0.0 else int(math.floor(math.log2(1024.0 / amax)))
    kk = max(min(kk, 24), -24)
    ws = (wp * (2.0 ** kk)).reshape(J * ch, k * 2 * 64)
    hi = ws.half()
    lo = (ws - hi.float()).half()
    return ToeplitzWeights(hi.contiguous(), lo.contiguous(), None if bias is None else bias.float().contiguous(), ch, cin, k,
                           pad, kk)


def conv_ps_weights(w: torch.Tensor, bias) -> ConvWeights:
    """Conv2d(Cin -> 4C, 3x3, pad 1) followed by PixelShuffle(2) (wxformer/crossformer.py:143-159, 813-830).

    out[c, 2y+dy, 2x+dx] = conv[4c + 2dy + dx, y, x]: phase z = 2dy + dx is a 3x3 conv with the weight rows 4c + z,
    scattered to (2y+dy, 2x+dx); the bias is re-ordered to [phase, C].
    """
    c4, cin, kh, kw = w.shape
    c = c4 // 4
    wz = w.reshape(c, 4, cin, kh, kw).permute(1, 0, 3, 4, 2).reshape(4, c, kh * kw * cin).contiguous()  # [z, c, (ky,kx,ci)]
    ky, kx = torch.meshgrid(torch.arange(kh), torch.arange(kw), indexing="ij")
    t1 = torch.stack([ky.reshape(-1) - 1, kx.reshape(-1) - 1], dim=-1).to(torch.int32)
    taps = t1.reshape(1, kh * kw, 2).repeat(4, 1, 1).contiguous().to(w.device)
    bz = bias.float().reshape(c, 4).t().contiguous().reshape(-1)
    return ConvWeights(wz, taps, bz, c, kh * kw, cin, 1, 4, 2, c)


@dataclass
class AttentionWeights:
    ln_g: torch.Tensor
    ln_b: torch.Tensor
    qkv: ConvWeights
    out: ConvWeights
    bias_t: torch.Tensor  # [L, L] transposed position bias
    wsz: int
    kind: int
    qkv_tc: Optional[GemmWeights] = None
    out_tc: Optional[GemmWeights] = None


@dataclass
class FeedForwardWeights:
    ln_g: torch.Tensor
    ln_b: torch.Tensor
    fc1: ConvWeights
    fc2: ConvWeights
    fc1_tc: Optional[GemmWeights] = None
    fc2_tc: Optional[GemmWeights] = None


@dataclass
class UpBlockWeights:
    up: ConvWeights
    convs: List[ConvWeights]
    gn_w: List[torch.Tensor]
    gn_b: List[torch.Tensor]
    up_tc: Optional[ConvTcWeights] = None
    convs_tc: Optional[List[ConvTcWeights]] = None
    sharp: Optional[ConvWeights] = None          # wxformer variant: sharpening conv after the PixelShuffle
    sharp_tc: Optional[ConvTcWeights] = None


@dataclass
class PreparedWeights:
    embeds: List[List[ConvWeights]]
    blocks: List[List[tuple]]  # per stage, per layer: (short_attn, ff, long_attn, ff)
    ups: List[UpBlockWeights]
    head: ConvWeights
    cin0_pad: int
    embeds_tc: Optional[List[List[Optional[ConvTcWeights]]]] = None
    head_tc: Optional[ConvTcWeights] = None
    embed0_toep: Optional[List[ToeplitzWeights]] = None  # None unless every stage-0 branch is eligible
    head2: Optional[ConvWeights] = None           # wxformer variant: conv3x3 after up_block4's PixelShuffle
    head2_tc: Optional[ConvTcWeights] = None


def prepare(sd: Dict[str, torch.Tensor], geo: Geometry, cin0_pad: int) -> PreparedWeights:
    """Fold and re-lay every parameter of the state dict for the kernels (device of ``sd``)."""
    wx = geo.variant == "wxformer"
    with torch.no_grad():
        embeds, blocks = [], []
        for st in geo.stages:
            s = st.index
            brs = []
            for i, br in enumerate(st.branches):
                ck = f"layers.{s}.0.convs.{i}" + (".1" if wx else "")
                w = fold_spectral_norm(sd, ck)
                brs.append(conv_weights(w, sd[ck + ".bias"], br.stride, br.pad, cin0_pad if s == 0 else None))
            embeds.append(brs)
            layers = []
            for l in range(st.depth):
                entry = []
                for a, kind, wsz in ((0, 0, st.local_window), (2, 1, st.global_window)):
                    p = f"layers.{s}.1.layers.{l}.{a}"
                    w_qkv, w_out = fold_spectral_norm(sd, p + ".to_qkv"), fold_spectral_norm(sd, p + ".to_out")
                    att = AttentionWeights(
                        sd[p + ".norm.g"].float().reshape(-1).contiguous(), sd[p + ".norm.b"].float().reshape(-1).contiguous(),
                        conv_weights(w_qkv, None, 1, 0),
                        conv_weights(w_out, sd[p + ".to_out.bias"], 1, 0),
                        position_bias_table(sd, p, wsz).t().contiguous(), wsz, kind,
                        gemm_weights(w_qkv, None), gemm_weights(w_out, sd[p + ".to_out.bias"]))
                    f = f"layers.{s}.1.layers.{l}.{a + 1}.layers"
                    w1, w2 = fold_spectral_norm(sd, f + ".1"), fold_spectral_norm(sd, f + ".4")
                    ff = FeedForwardWeights(
                        sd[f + ".0.g"].float().reshape(-1).contiguous(), sd[f + ".0.b"].float().reshape(-1).contiguous(),
                        conv_weights(w1, sd[f + ".1.bias"], 1, 0),
                        conv_weights(w2, sd[f + ".4.bias"], 1, 0),
                        gemm_weights(w1, sd[f + ".1.bias"]), gemm_weights(w2, sd[f + ".4.bias"]))
                    entry += [att, ff]
                layers.append(tuple(entry))
            blocks.append(layers)
        ups = []
        for up in geo.ups:
            n = up.name
            if wx:
                upw = conv_ps_weights(fold_spectral_norm(sd, n + ".conv"), sd[n + ".conv.bias"])
            else:
                upw = convt_k2s2_weights(fold_spectral_norm(sd, n + ".conv", 1), sd[n + ".conv.bias"])
            ups.append(UpBlockWeights(
                upw,
                [conv_weights(fold_spectral_norm(sd, f"{n}.b.{ci}"), sd[f"{n}.b.{ci}.bias"], 1, 1) for ci in (0, 3)],
                [sd[f"{n}.b.{gi}.weight"].float().contiguous() for gi in (1, 4)],
                [sd[f"{n}.b.{gi}.bias"].float().contiguous() for gi in (1, 4)]))
            if wx:
                ups[-1].sharp = conv_weights(fold_spectral_norm(sd, n + ".sharp"), sd[n + ".sharp.bias"], 1, 1)
        head2 = None
        if wx:
            head = conv_ps_weights(fold_spectral_norm(sd, "up_block4.0"), sd["up_block4.0.bias"])
            head2 = conv_weights(fold_spectral_norm(sd, "up_block4.2"), sd["up_block4.2.bias"], 1, 1)
        else:
            head = convt_k4s2p1_weights(fold_spectral_norm(sd, "up_block4", 1), sd["up_block4.bias"])
        # tensor-core planes for every convolution whose output-channel count suits the epilogue (N % 4 == 0)
        for uw in ups:
            uw.up_tc = conv_tc_weights(uw.up)
            uw.convs_tc = [conv_tc_weights(c) for c in uw.convs]
            if uw.sharp is not None:
                uw.sharp_tc = conv_tc_weights(uw.sharp)
        embeds_tc = [[(conv_tc_weights(c) if (s > 0 and c.n % 4 == 0 and c.t <= 64) else None) for c in brs]
                     for s, brs in enumerate(embeds)]
        head_tc = conv_tc_weights(head) if head.n % 4 == 0 else None
        head2_tc = conv_tc_weights(head2) if (head2 is not None and head2.n % 8 == 0) else None  # its input rows are TMA rows
        st0 = geo.stages[0]
        toep = None
        if all(toeplitz_eligible(br.c_out, st0.c_in, br.kernel, br.stride) for br in st0.branches):
            sfx = ".1" if wx else ""
            toep = [toeplitz_weights(fold_spectral_norm(sd, f"layers.0.0.convs.{i}{sfx}"),
                                     sd[f"layers.0.0.convs.{i}{sfx}.bias"], br.pad) for i, br in enumerate(st0.branches)]
    return PreparedWeights(embeds, blocks, ups, head, cin0_pad, embeds_tc, head_tc, toep, head2, head2_tc)
