"""ORACLE — test infrastructure only.  Parity status: PINNED (see tests/golden/README.md).

A CPU restatement, in plain fp32 PyTorch tensor ops, of the reference's forecast forward step
``y = CrossFormer(x)`` (``/root/reference/credit/models/crossformer.py:593-644``), written as
pure functions of a *state dict* (no nn.Module tree) so that it shares no structure with the
thing it checks.  Every function cites the reference lines it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this file.  The product (``miles_credit_b200``) never does: it runs hand-written
sm_100a kernels through the C-ABI in ``include/wxformer_b200.h`` and fails loudly without them.

Pinning: ``tests/golden/make_golden.py`` (run in the build container, where ``/root/reference`` is
mounted) imports the UNMODIFIED reference module through ``credit.models.load_model``, loads the
same synthetic state dict, and stores the reference's outputs and per-block activations in
``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` checks this file against those vectors.
"""

from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from miles_credit_b200.geometry import Geometry


# --------------------------------------------------------------------------------------
# boundary padding (credit/boundary_padding.py)


def earth_pad_index_map(h: int, w: int, pad_lat, pad_lon):
    """Source (row, col) of every padded pixel, as plain integer arithmetic.

    Follows ``TensorPadding._earth_padding`` (boundary_padding.py:50-72): the pole rows come from
    the 180-degree-rolled field flipped in latitude (the pole row itself is repeated), then the
    lat-padded field is wrapped circularly in longitude.
    """
    pt, pb = pad_lat
    pl, pr = pad_lon
    hp, wp = h + pt + pb, w + pl + pr
    rows = torch.empty(hp, dtype=torch.long)
    flip = torch.zeros(hp, dtype=torch.bool)
    for r in range(hp):
        if r < pt:
            rows[r], flip[r] = pt - 1 - r, True
        elif r < pt + h:
            rows[r] = r - pt
        else:
            rows[r], flip[r] = h - 1 - (r - pt - h), True
    j = (torch.arange(wp) - pl) % w
    shift = w // 2
    cols_plain = j
    cols_roll = (j - shift) % w
    return rows, flip, cols_plain, cols_roll


def pad_field(x: torch.Tensor, mode: str, pad_lat, pad_lon) -> torch.Tensor:
    """``TensorPadding.pad`` on [..., H, W] (boundary_padding.py:20-33, 50-72, 98-117)."""
    h, w = x.shape[-2:]
    pt, pb = pad_lat
    pl, pr = pad_lon
    if mode == "earth":
        rows, flip, cp, cr = earth_pad_index_map(h, w, pad_lat if (pt > 0 or pb > 0) else (0, 0), pad_lon)
        if not (pl > 0 or pr > 0):
            cp, cr = torch.arange(w), (torch.arange(w) - w // 2) % w
        cols = torch.where(flip[:, None], cr[None, :], cp[None, :])  # [Hp, Wp]
        return x[..., rows[:, None], cols]
    if mode == "mirror":
        # circular in longitude first, then reflect (no edge repeat) in latitude
        wp = w + pl + pr
        cols = (torch.arange(wp) - pl) % w if (pl > 0 or pr > 0) else torch.arange(w)
        hp = h + pt + pb if (pt > 0 or pb > 0) else h
        r = torch.arange(hp) - (pt if (pt > 0 or pb > 0) else 0)
        r = torch.where(r < 0, -r, r)
        r = torch.where(r >= h, 2 * (h - 1) - r, r)
        return x[..., r[:, None], cols[None, :]]
    raise ValueError(mode)


def unpad_field(x: torch.Tensor, pad_lat, pad_lon) -> torch.Tensor:
    """``TensorPadding.unpad`` (boundary_padding.py:74-96, 119-137): plain crop."""
    pt, pb = pad_lat
    pl, pr = pad_lon
    h, w = x.shape[-2:]
    if pt > 0 or pb > 0:
        x = x[..., pt : h - pb, :]
    if pl > 0 or pr > 0:
        x = x[..., :, pl : w - pr]
    return x


# --------------------------------------------------------------------------------------
# weights


def effective_weight(sd: Dict[str, torch.Tensor], prefix: str, sn_dim: int = 0) -> torch.Tensor:
    """Eval-mode weight of a (possibly spectral-normed) module.

    torch.nn.utils.spectral_norm in eval(): ``W = weight_orig / (u . (W_mat v))`` with the stored
    ``weight_u/weight_v`` and no power iteration (hook applied at crossformer.py:23-26, 576-578).
    """
    if prefix + ".weight_orig" not in sd:
        return sd[prefix + ".weight"]
    w = sd[prefix + ".weight_orig"]
    wm = w
    if sn_dim != 0:
        wm = w.permute(sn_dim, *[d for d in range(w.dim()) if d != sn_dim])
    wm = wm.reshape(wm.shape[0], -1)
    sigma = torch.dot(sd[prefix + ".weight_u"], torch.mv(wm, sd[prefix + ".weight_v"]))
    return w / sigma


def position_bias(sd, prefix: str, wsz: int) -> torch.Tensor:
    """[L, L] relative-position bias of one Attention (crossformer.py:158-176, 238-245, 279-286).

    The index buffer uses row stride (2w-1) into an MLP table laid out with row stride (2w+1);
    that quirk is part of the reference's arithmetic and is kept as is.
    """
    pos = torch.arange(-wsz, wsz + 1, dtype=torch.float32)
    gy, gx = torch.meshgrid(pos, pos, indexing="ij")
    t = torch.stack([gy.reshape(-1), gx.reshape(-1)], dim=-1)  # [(2w+1)^2, 2]
    for lin, ln in ((0, 1), (3, 4), (6, 7)):
        t = F.linear(t, effective_weight(sd, f"{prefix}.dpb.layers.{lin}"), sd[f"{prefix}.dpb.layers.{lin}.bias"])
        t = F.layer_norm(t, (t.shape[-1],), sd[f"{prefix}.dpb.layers.{ln}.weight"],
                         sd[f"{prefix}.dpb.layers.{ln}.bias"], 1e-5)
        t = torch.relu(t)
    t = F.linear(t, effective_weight(sd, f"{prefix}.dpb.layers.9"), sd[f"{prefix}.dpb.layers.9.bias"]).squeeze(-1)
    p = torch.arange(wsz)
    ty, tx = torch.meshgrid(p, p, indexing="ij")
    tok = torch.stack([ty.reshape(-1), tx.reshape(-1)], dim=-1)  # token (row, col) inside a window
    rel = tok[:, None, :] - tok[None, :, :] + (wsz - 1)
    idx = rel[..., 0] * (2 * wsz - 1) + rel[..., 1]
    return t[idx]


# --------------------------------------------------------------------------------------
# encoder blocks


def channel_layer_norm(x, g, b, eps: float = 1e-5):
    """Custom channel LayerNorm on NCHW (crossformer.py:182-192): biased variance, eps inside sqrt."""
    var = x.var(dim=1, unbiased=False, keepdim=True)
    mean = x.mean(dim=1, keepdim=True)
    return (x - mean) / (var + eps).sqrt() * g + b


def cross_embed(x, sd, prefix: str, stage) -> torch.Tensor:
    """CrossEmbedLayer.forward (crossformer.py:128-152): concat of strided convs, sorted kernels."""
    outs = []
    for i, br in enumerate(stage.branches):
        if f"{prefix}.convs.{i}.1.bias" in sd:
            # wxformer variant (wxformer/crossformer.py:199-236): ZeroPad2d(left=(k-s)//2, right=(k-s)-left) + unpadded conv
            key = f"{prefix}.convs.{i}.1"
            tot = br.kernel - br.stride
            lo, hi = tot // 2, tot - tot // 2
            outs.append(F.conv2d(F.pad(x, (lo, hi, lo, hi)), effective_weight(sd, key), sd[key + ".bias"], stride=br.stride))
        else:
            w = effective_weight(sd, f"{prefix}.convs.{i}")
            outs.append(F.conv2d(x, w, sd[f"{prefix}.convs.{i}.bias"], stride=br.stride, padding=br.pad))
    return torch.cat(outs, dim=1)


def window_attention(x, sd, prefix: str, kind: str, wsz: int, heads: int, dim_head: int) -> torch.Tensor:
    """Attention.forward (crossformer.py:247-316) without the residual.

    short: contiguous wsz x wsz tiles.  long: token (l1, l2) of group (gh, gw) sits at row
    l1*(H/wsz)+gh, col l2*(W/wsz)+gw.
    """
    b, d, hh, ww = x.shape
    xn = channel_layer_norm(x, sd[prefix + ".norm.g"], sd[prefix + ".norm.b"])
    nh, nw = hh // wsz, ww // wsz
    if kind == "short":
        t = xn.reshape(b, d, nh, wsz, nw, wsz).permute(0, 2, 4, 1, 3, 5)  # b gh gw d s1 s2
    else:
        t = xn.reshape(b, d, wsz, nh, wsz, nw).permute(0, 3, 5, 1, 2, 4)  # b gh gw d l1 l2
    t = t.reshape(b * nh * nw, d, wsz * wsz)  # windows, channels, tokens
    wqkv = effective_weight(sd, prefix + ".to_qkv").reshape(3 * d, d)
    qkv = torch.einsum("od,ndl->nol", wqkv, t)
    q, k, v = qkv.split(d, dim=1)
    nwin, L = t.shape[0], wsz * wsz

    def heads_of(z):
        return z.reshape(nwin, heads, dim_head, L).transpose(2, 3)  # n h L dh

    q, k, v = heads_of(q) * (dim_head**-0.5), heads_of(k), heads_of(v)
    sim = q @ k.transpose(-1, -2) + position_bias(sd, prefix, wsz)
    attn = sim.softmax(dim=-1)
    o = (attn @ v).transpose(2, 3).reshape(nwin, d, L)
    wo = effective_weight(sd, prefix + ".to_out").reshape(d, d)
    o = torch.einsum("od,ndl->nol", wo, o) + sd[prefix + ".to_out.bias"][None, :, None]
    if kind == "short":
        o = o.reshape(b, nh, nw, d, wsz, wsz).permute(0, 3, 1, 4, 2, 5)
    else:
        o = o.reshape(b, nh, nw, d, wsz, wsz).permute(0, 3, 4, 1, 5, 2)
    return o.reshape(b, d, hh, ww)


def feed_forward(x, sd, prefix: str) -> torch.Tensor:
    """FeedForward.forward (crossformer.py:195-207): LN, 1x1 d->4d, exact-erf GELU, 1x1 4d->d."""
    xn = channel_layer_norm(x, sd[prefix + ".layers.0.g"], sd[prefix + ".layers.0.b"])
    h = F.conv2d(xn, effective_weight(sd, prefix + ".layers.1"), sd[prefix + ".layers.1.bias"])
    h = F.gelu(h)
    return F.conv2d(h, effective_weight(sd, prefix + ".layers.4"), sd[prefix + ".layers.4.bias"])


def transformer_stage(x, sd, stage, dim_head: int, taps: Optional[dict] = None) -> torch.Tensor:
    """Transformer.forward (crossformer.py:358-365): (short attn, FF, long attn, FF) x depth, all residual."""
    s = stage.index
    for l in range(stage.depth):
        p = f"layers.{s}.1.layers.{l}"
        x = window_attention(x, sd, p + ".0", "short", stage.local_window, stage.heads, dim_head) + x
        if taps is not None:
            taps[f"s{s}.l{l}.short_attn"] = x
        x = feed_forward(x, sd, p + ".1") + x
        x = window_attention(x, sd, p + ".2", "long", stage.global_window, stage.heads, dim_head) + x
        if taps is not None:
            taps[f"s{s}.l{l}.long_attn"] = x
        x = feed_forward(x, sd, p + ".3") + x
    return x


# --------------------------------------------------------------------------------------
# decoder


def up_block(x, sd, up) -> torch.Tensor:
    """UpBlock.forward (crossformer.py:107-122) with attention=None: ConvT k2s2, 2x(conv3x3, GN, SiLU), + skip."""
    n = up.name
    x = F.conv_transpose2d(x, effective_weight(sd, n + ".conv", sn_dim=1), sd[n + ".conv.bias"], stride=2)
    y = x
    for ci, gi in ((0, 1), (3, 4)):
        y = F.conv2d(y, effective_weight(sd, f"{n}.b.{ci}"), sd[f"{n}.b.{ci}.bias"], padding=1)
        y = F.group_norm(y, up.groups, sd[f"{n}.b.{gi}.weight"], sd[f"{n}.b.{gi}.bias"], 1e-5)
        y = F.silu(y)
    return y + x


def pixel_shuffle2(x: torch.Tensor) -> torch.Tensor:
    """nn.PixelShuffle(2) spelled out: out[c, 2y+dy, 2x+dx] = in[4c + 2dy + dx, y, x]."""
    b, c4, h, w = x.shape
    c = c4 // 4
    return x.reshape(b, c, 2, 2, h, w).permute(0, 1, 4, 2, 5, 3).reshape(b, c, 2 * h, 2 * w)


def up_block_ps(x, sd, up) -> torch.Tensor:
    """UpBlockPS.forward (wxformer/crossformer.py:156-162): conv3x3 -> PixelShuffle -> x + sharp(x) -> stack + skip."""
    n = up.name
    x = pixel_shuffle2(F.conv2d(x, effective_weight(sd, n + ".conv"), sd[n + ".conv.bias"], padding=1))
    x = x + F.conv2d(x, effective_weight(sd, n + ".sharp"), sd[n + ".sharp.bias"], padding=1)
    y = x
    for ci, gi in ((0, 1), (3, 4)):
        y = F.conv2d(y, effective_weight(sd, f"{n}.b.{ci}"), sd[f"{n}.b.{ci}.bias"], padding=1)
        y = F.group_norm(y, up.groups, sd[f"{n}.b.{gi}.weight"], sd[f"{n}.b.{gi}.bias"], 1e-5)
        y = F.silu(y)
    return y + x


def bilinear_resize(x, h_out: int, w_out: int) -> torch.Tensor:
    """``F.interpolate(mode="bilinear")`` with align_corners=False (crossformer.py:631-632), spelled out.

    src = (dst + 0.5) * in/out - 0.5 clamped at 0; neighbours clamped at the last row/col.
    """
    h_in, w_in = x.shape[-2:]

    def axis(n_in, n_out):
        # torch evaluates scale*(dst+0.5)-0.5 in fp32 with a fused multiply-add (one rounding), on CPU builds
        # and in the CUDA kernel alike; at 720->721 an unfused product moves the weights by ~3e-5.
        scale = (torch.tensor(float(n_in), dtype=torch.float32) / torch.tensor(float(n_out), dtype=torch.float32)).double()
        src = (scale * (torch.arange(n_out, dtype=torch.float64) + 0.5) - 0.5).float()
        src = src.clamp_min(0.0)
        i0 = src.floor().long().clamp_max(n_in - 1)
        i1 = (i0 + 1).clamp_max(n_in - 1)
        lam = (src - i0.float()).clamp(0.0, 1.0)
        return i0, i1, lam

    y0, y1, ly = axis(h_in, h_out)
    x0, x1, lx = axis(w_in, w_out)
    top = x[..., y0, :][..., x0] * (1 - lx) + x[..., y0, :][..., x1] * lx
    bot = x[..., y1, :][..., x0] * (1 - lx) + x[..., y1, :][..., x1] * lx
    return top * (1 - ly)[:, None] + bot * ly[:, None]


# --------------------------------------------------------------------------------------
# the whole step


def forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], geo: Geometry, taps: Optional[dict] = None) -> torch.Tensor:
    """CrossFormer.forward (crossformer.py:593-644) for patch=1, post_conf off.

    x: [B, C_in, frames, H, W] fp32  ->  [B, C_out, output_frames, H_out, W_out] fp32.
    ``taps`` (optional dict) receives named intermediate activations for block-level parity tests.
    """
    b = x.shape[0]
    if geo.padding.activate:
        x = pad_field(x, geo.padding.mode, geo.padding.pad_lat, geo.padding.pad_lon)
    if taps is not None:
        taps["padded"] = x
    # frames flattened channel-major, time-minor (crossformer.py:604-609)
    x = x.reshape(b, geo.input_channels, x.shape[-2], x.shape[-1])
    enc: List[torch.Tensor] = []
    for st in geo.stages:
        x = cross_embed(x, sd, f"layers.{st.index}.0", st)
        if taps is not None:
            taps[f"s{st.index}.embed"] = x
        x = transformer_stage(x, sd, st, geo.dim_head, taps)
        if taps is not None:
            taps[f"s{st.index}.out"] = x
        enc.append(x)
    wx = geo.variant == "wxformer"
    for up, skip in zip(geo.ups, (2, 1, 0)):
        x = up_block_ps(x, sd, up) if wx else up_block(x, sd, up)
        if taps is not None:
            taps[up.name] = x
        x = torch.cat([x, enc[skip]], dim=1)
    if wx:  # up_block4 = conv3x3 -> PixelShuffle(2) -> conv3x3 (wxformer/crossformer.py:813-830)
        x = pixel_shuffle2(F.conv2d(x, effective_weight(sd, "up_block4.0"), sd["up_block4.0.bias"], padding=1))
        x = F.conv2d(x, effective_weight(sd, "up_block4.2"), sd["up_block4.2.bias"], padding=1)
    else:
        x = F.conv_transpose2d(x, effective_weight(sd, "up_block4", sn_dim=1), sd["up_block4.bias"], stride=2, padding=1)
    if taps is not None:
        taps["up_block4"] = x
    if geo.padding.activate:
        x = unpad_field(x, geo.padding.pad_lat, geo.padding.pad_lon)
    if geo.interp:
        x = bilinear_resize(x, geo.image_height, geo.image_width)
    return x.reshape(b, geo.base_output_channels, geo.output_frames, x.shape[-2], x.shape[-1])


def rollout_update(x: torch.Tensor, y: torch.Tensor, geo: Geometry, forcing: Optional[torch.Tensor] = None):
    """One autoregressive state update for frames == 1 (``update_x``, datasets/gen_2/channel_utils.py:253-291).

    Prognostic channels (levels*channels + surface) of the next input come from the prediction,
    dynamic-forcing/static channels are replaced by ``forcing`` if given, else carried over.
    """
    n_prog = geo.channels * geo.levels + geo.surface_channels
    nxt = x.clone()
    nxt[:, :n_prog, -1] = y[:, :n_prog, 0]
    if forcing is not None:
        nxt[:, n_prog:, -1] = forcing
    return nxt
