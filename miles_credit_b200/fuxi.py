"""``FuxiB200`` — drop-in replacement of ``credit.models.fuxi.Fuxi`` (registry key ``fuxi``) for the forecast step.

Same constructor keywords (fuxi.py:321-352), same ``state_dict`` keys and shapes (fuxi.py's own modules plus the
parameter tree of timm's ``SwinTransformerV2Stage`` it instantiates, fuxi.py:250-260), same tensor contract
``[B, C_in, frames, H, W] -> [B, C_out, 1, H, W]`` (fuxi.py:454-506).  The forward is a fixed launch plan over the
sm_100a kernels of the C ABI (include/wxformer_b200.h):

  pad + frame fold (wxf_pad_to_pixel_major_f16x2)  ->  CubeEmbedding: Conv3d k = stride = (T, ph, pw) as a 4x4 stride-4
  implicit GEMM + LayerNorm (fuxi.py:82-143)  ->  DownBlock: conv3x3 s2, 2 x (conv3x3 + GroupNorm + SiLU) + skip (:146-172)
  ->  zero pad to a window multiple (:67-79, 281-283) as a row gather  ->  depth x Swin-V2 block: qkv GEMM, scaled-cosine
  window attention (shift, masks, per-head bias / scale), proj GEMM, x + norm1(.), fc1 + GELU, fc2, x + norm2(.)
  ->  crop + concat with the shortcut as a row gather (:288-292)  ->  UpBlock: ConvT k2 s2, residual stack (:175-201)
  ->  dense head GEMM (:420, 484)  ->  un-patchify + un-pad + bilinear + NCHW (:485-498).

Eval-mode forward only; no CPU path, no PyTorch fallback.  The Swin-V2 stage is third-party ``timm`` code that is absent
from the reference tree and this image: its arithmetic here follows ``oracle/swin_v2.py`` (pinned bit-exactly against
HuggingFace's independent ``Swinv2Stage`` port, tests/test_swin_v2_vs_hf.py; against timm itself: unpinned, SURVEY.md section
8c); everything FuXi owns in the reference tree is pinned by ``tests/golden/unit_fuxi*.pt``.
One reference quirk is kept: under FuXi's old-style spectral-norm hooks timm's qkv projection runs on the UN-normalised
``weight_orig`` (timm calls ``F.linear(x, self.qkv.weight, ...)``, so the hook never fires).
"""

from __future__ import annotations

import logging
import math
import os
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import lib as _lib
from . import ops
from .geometry import Padding, _sn
from .model import PaddingView, _Base, _Holder, _round_up, _to_device
from .synth import synthesize
from .weights import (ConvTcWeights, GemmWeights, conv_tc_weights, conv_weights, convt_k2s2_weights, fold_spectral_norm,
                      gemm_weights)

logger = logging.getLogger(__name__)


def _pad_to_window(n: int, w: int) -> Tuple[int, int]:
    """(before, after) zero padding to a multiple of the window, smaller half first (get_pad2d, fuxi.py:25-79)."""
    if n % w == 0:
        return 0, 0
    p = w - n % w
    return p // 2, p - p // 2


@dataclass(frozen=True)
class FuxiGeometry:
    image_height: int
    image_width: int
    patch_height: int
    patch_width: int
    frames: int
    levels: int
    channels: int
    surface_channels: int
    input_only_channels: int
    output_only_channels: int
    in_chans: int
    out_chans: int
    dim: int
    num_groups: int
    num_heads: int
    depth: int
    window_size: int
    use_spectral_norm: bool
    interp: bool
    padding: Padding
    h_pad: int
    w_pad: int
    lat: int      # patch grid (CubeEmbedding output)
    lon: int
    th: int       # token grid of the Swin stage = patch grid / 2 (fuxi.py:402-405)
    tw: int
    pad2d: Tuple[int, int, int, int]  # (left, right, top, bottom) zero padding of the token grid
    gh: int       # padded token grid
    gw: int
    ws: Tuple[int, int]      # window, clamped to the grid (timm _calc_window_shift)
    shift: Tuple[int, int]   # shift of the odd blocks

    @property
    def in_shape(self):
        return (self.in_chans, self.frames, self.image_height, self.image_width)

    @property
    def out_shape(self):
        return (self.out_chans, 1, self.h_out, self.w_out)

    @property
    def h_crop(self):
        return self.h_pad - sum(self.padding.pad_lat)

    @property
    def w_crop(self):
        return self.w_pad - sum(self.padding.pad_lon)

    @property
    def h_out(self):
        return self.image_height if self.interp else self.h_crop

    @property
    def w_out(self):
        return self.image_width if self.interp else self.w_crop

    @property
    def dh(self):
        return self.dim // self.num_heads

    @property
    def output_frames(self):
        return 1

    def block_shift(self, i: int) -> Tuple[int, int]:
        return (0, 0) if i % 2 == 0 else self.shift


def build_fuxi_geometry(image_height=640, patch_height=16, image_width=1280, patch_width=16, levels=15, frames=2,
                        frame_patch_size=2, dim=1536, num_groups=32, channels=4, surface_channels=7, input_only_channels=0,
                        output_only_channels=0, num_heads=8, depth=48, window_size=7, use_spectral_norm=True, interp=True,
                        proj_drop=0, attn_drop=0, drop_path=0, padding_conf=None, post_conf=None, use_noise=False,
                        noise_latent_dim=128, noise_factor=0.2, noise_scheduler=None, freeze=False, **kwargs) -> FuxiGeometry:
    """Keyword surface and defaults of ``Fuxi.__init__`` (fuxi.py:321-352); unknown keys are swallowed like there."""
    if post_conf is not None and post_conf.get("activate", False):
        raise NotImplementedError("in-model PostBlock is out of scope; use post_conf.activate=False")
    if use_noise:
        raise NotImplementedError("use_noise=True cannot be constructed in the reference either (fuxi.py:267-272 passes a "
                                  "`scheduler` keyword StochasticDecompositionLayer does not take)")
    if proj_drop or attn_drop or (drop_path if not isinstance(drop_path, (list, tuple)) else any(drop_path)):
        raise NotImplementedError("dropout / drop-path are training features; the forecast step is eval-mode")
    if frames != frame_patch_size:
        raise NotImplementedError("frames != frame_patch_size leaves more than one time slice after the cube embedding; "
                                  "the reference's squeeze(2) (fuxi.py:475) then does nothing and the model fails")
    if patch_height != patch_width:
        raise NotImplementedError("the patch-embedding kernel takes one stride for both axes (patch_height == patch_width)")
    if patch_height not in (1, 2, 4):
        raise NotImplementedError("patch size must be 1, 2 or 4 (TMA element stride of the implicit-GEMM patch embedding)")
    padding = Padding.from_conf(padding_conf)
    in_chans = channels * levels + surface_channels + input_only_channels
    out_chans = channels * levels + surface_channels + output_only_channels
    h_pad = image_height + sum(padding.pad_lat)
    w_pad = image_width + sum(padding.pad_lon)
    if h_pad % patch_height or w_pad % patch_width:
        raise ValueError(f"padded grid {h_pad}x{w_pad} is not a multiple of the patch {patch_height}x{patch_width}")
    lat, lon = h_pad // patch_height, w_pad // patch_width
    if lat % 2 or lon % 2:
        raise ValueError(f"patch grid {lat}x{lon} must be even: DownBlock halves it and UpBlock doubles it back (fuxi.py:146-201)")
    th, tw = round(h_pad / patch_height / 2), round(w_pad / patch_width / 2)
    top, bottom = _pad_to_window(th, window_size)
    left, right = _pad_to_window(tw, window_size)
    gh, gw = th + top + bottom, tw + left + right
    ws = tuple(r if r <= window_size else window_size for r in (gh, gw))
    shift = tuple(0 if r <= w else window_size // 2 for r, w in zip((gh, gw), ws))
    if dim % num_heads or (dim // num_heads) % 4:
        raise ValueError(f"dim {dim} / heads {num_heads}: the head dimension must be a multiple of 4")
    if dim % num_groups:
        raise ValueError(f"GroupNorm: {dim} channels not divisible by {num_groups} groups")
    if ws[0] * ws[1] > 64:
        raise NotImplementedError(f"window {ws} holds {ws[0] * ws[1]} tokens; the attention kernel holds at most 64")
    if dim % 8:
        raise NotImplementedError("dim must be a multiple of 8 (16-byte operand rows)")
    return FuxiGeometry(image_height, image_width, patch_height, patch_width, frames, levels, channels, surface_channels,
                        input_only_channels, output_only_channels, in_chans, out_chans, dim, num_groups, num_heads, depth,
                        window_size, bool(use_spectral_norm), bool(interp), padding, h_pad, w_pad, lat, lon, th, tw,
                        (left, right, top, bottom), gh, gw, ws, shift)


def fuxi_state_spec(geo: FuxiGeometry) -> "OrderedDict[str, Tuple[Tuple[int, ...], str]]":
    """Every persistent tensor of the reference module (fuxi.py + timm's stage): key -> (shape, role)."""
    sn, d = geo.use_spectral_norm, geo.dim
    spec: "OrderedDict[str, Tuple[Tuple[int, ...], str]]" = OrderedDict()
    # Conv3d is skipped by apply_spectral_norm (fuxi.py:17-23)
    spec["cube_embedding.proj.weight"] = ((d, geo.in_chans, geo.frames, geo.patch_height, geo.patch_width), "weight:0")
    spec["cube_embedding.proj.bias"] = ((d,), "bias")
    spec["cube_embedding.norm.weight"] = ((d,), "gain")
    spec["cube_embedding.norm.bias"] = ((d,), "shift")

    def stack(prefix):
        for ci, gi in ((0, 1), (3, 4)):
            _sn(spec, f"{prefix}.b.{ci}", (d, d, 3, 3), sn)
            spec[f"{prefix}.b.{gi}.weight"] = ((d,), "gain")
            spec[f"{prefix}.b.{gi}.bias"] = ((d,), "shift")

    _sn(spec, "u_transformer.down.conv", (d, d, 3, 3), sn)
    stack("u_transformer.down")
    for i in range(geo.depth):
        p = f"u_transformer.layer.blocks.{i}"
        spec[p + ".attn.logit_scale"] = ((geo.num_heads, 1, 1), f"const:{math.log(10.0)}")
        spec[p + ".attn.q_bias"] = ((d,), "bias")
        spec[p + ".attn.v_bias"] = ((d,), "bias")
        _sn(spec, p + ".attn.cpb_mlp.0", (512, 2), sn)
        _sn(spec, p + ".attn.cpb_mlp.2", (geo.num_heads, 512), sn, bias=False)
        _sn(spec, p + ".attn.qkv", (3 * d, d), sn, bias=False)
        _sn(spec, p + ".attn.proj", (d, d), sn)
        spec[p + ".norm1.weight"] = ((d,), "gain")
        spec[p + ".norm1.bias"] = ((d,), "shift")
        _sn(spec, p + ".mlp.fc1", (4 * d, d), sn)
        _sn(spec, p + ".mlp.fc2", (d, 4 * d), sn)
        spec[p + ".norm2.weight"] = ((d,), "gain")
        spec[p + ".norm2.bias"] = ((d,), "shift")
    _sn(spec, "u_transformer.up.conv", (2 * d, d, 2, 2), sn, sn_dim=1, bias_len=d)
    stack("u_transformer.up")
    _sn(spec, "fc", (geo.out_chans * geo.patch_height * geo.patch_width, d), sn)
    return spec


def synthetic_fuxi_state_dict(geo: FuxiGeometry, seed: int = 1000, sn_iters: int = 5):
    """Deterministic synthetic weights (no network for checkpoints): the WXFormer recipe (synth.py) on FuXi's key table."""
    return synthesize(fuxi_state_spec(geo), seed, sn_iters)


def synthetic_fuxi_input(geo: FuxiGeometry, batch: int = 1, seed: int = 1000) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed * 7919 + 17)
    return torch.randn((batch, *geo.in_shape), generator=g, dtype=torch.float32)


# ---- load-time weight preparation ---------------------------------------------------------------------------------

def _rel_coords_table(ws):
    ch = torch.arange(-(ws[0] - 1), ws[0], dtype=torch.float32)
    cw = torch.arange(-(ws[1] - 1), ws[1], dtype=torch.float32)
    table = torch.stack(torch.meshgrid(ch, cw, indexing="ij")).permute(1, 2, 0).contiguous()
    table[..., 0] /= max(ws[0] - 1, 1)
    table[..., 1] /= max(ws[1] - 1, 1)
    table = table * 8
    return torch.sign(table) * torch.log2(table.abs() + 1.0) / math.log2(8)


def _rel_index(ws):
    coords = torch.stack(torch.meshgrid(torch.arange(ws[0]), torch.arange(ws[1]), indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws[0] - 1
    rel[:, :, 1] += ws[1] - 1
    rel[:, :, 0] *= 2 * ws[1] - 1
    return rel.sum(-1)


def swin_position_bias(sd, prefix: str, ws, heads: int) -> torch.Tensor:
    """[heads, L, L] = 16 * sigmoid(cpb_mlp(log-spaced relative coordinates)) (timm WindowAttention, Swin-V2): the
    continuous position bias is input independent, so it is evaluated once per load."""
    t = _rel_coords_table(ws).to(sd[prefix + ".cpb_mlp.0.bias"].device)
    h = torch.relu(F.linear(t, fold_spectral_norm(sd, prefix + ".cpb_mlp.0"), sd[prefix + ".cpb_mlp.0.bias"].float()))
    table = F.linear(h, fold_spectral_norm(sd, prefix + ".cpb_mlp.2")).view(-1, heads)
    n = ws[0] * ws[1]
    rpb = table[_rel_index(ws).view(-1).to(table.device)].view(n, n, heads).permute(2, 0, 1)
    return (16 * torch.sigmoid(rpb)).contiguous()


@dataclass
class SwinBlockWeights:
    qkv: GemmWeights
    proj: GemmWeights
    fc1: GemmWeights
    fc2: GemmWeights
    bias: torch.Tensor         # [heads, L, L]
    logit_scale: torch.Tensor  # [heads] = exp(min(logit_scale, ln 100))
    n1_g: torch.Tensor
    n1_b: torch.Tensor
    n2_g: torch.Tensor
    n2_b: torch.Tensor


@dataclass
class StackWeights:
    convs: List[ConvTcWeights]
    gn_w: List[torch.Tensor]
    gn_b: List[torch.Tensor]


@dataclass
class FuxiWeights:
    cube: ConvTcWeights
    cube_g: torch.Tensor
    cube_b: torch.Tensor
    down: ConvTcWeights
    down_stack: StackWeights
    blocks: List[SwinBlockWeights]
    up: ConvTcWeights
    up_stack: StackWeights
    head: GemmWeights
    cp: int  # channels per pixel in the head's output columns (>= out_chans)


def prepare_fuxi(sd: Dict[str, torch.Tensor], geo: FuxiGeometry) -> FuxiWeights:
    d = geo.dim
    with torch.no_grad():
        w = sd["cube_embedding.proj.weight"].float()
        # Conv3d [d, C, T, ph, pw] with kernel = stride: the frames fold into the channels (index c*T + t, the order the
        # padding pass writes), leaving a 2-D ph x pw stride-ph convolution
        w2 = w.reshape(d, geo.in_chans * geo.frames, geo.patch_height, geo.patch_width)
        cube = conv_tc_weights(conv_weights(w2, sd["cube_embedding.proj.bias"], geo.patch_height, 0))

        def stack(prefix):
            return StackWeights(
                [conv_tc_weights(conv_weights(fold_spectral_norm(sd, f"{prefix}.b.{ci}"), sd[f"{prefix}.b.{ci}.bias"], 1, 1))
                 for ci in (0, 3)],
                [sd[f"{prefix}.b.{gi}.weight"].float().contiguous() for gi in (1, 4)],
                [sd[f"{prefix}.b.{gi}.bias"].float().contiguous() for gi in (1, 4)])

        n = "u_transformer"
        down = conv_tc_weights(conv_weights(fold_spectral_norm(sd, f"{n}.down.conv"), sd[f"{n}.down.conv.bias"], 2, 1))
        blocks = []
        for i in range(geo.depth):
            p = f"{n}.layer.blocks.{i}"
            # timm never calls the qkv module, so FuXi's spectral-norm hook never rewrites its weight: raw weight_orig
            qkv_w = sd.get(f"{p}.attn.qkv.weight_orig", sd.get(f"{p}.attn.qkv.weight")).float()
            qkv_b = torch.cat([sd[f"{p}.attn.q_bias"].float(), torch.zeros_like(sd[f"{p}.attn.v_bias"]).float(),
                               sd[f"{p}.attn.v_bias"].float()])
            ls = torch.clamp(sd[f"{p}.attn.logit_scale"].float().reshape(-1), max=math.log(1.0 / 0.01)).exp()
            blocks.append(SwinBlockWeights(
                gemm_weights(qkv_w, qkv_b),
                gemm_weights(fold_spectral_norm(sd, f"{p}.attn.proj"), sd[f"{p}.attn.proj.bias"]),
                gemm_weights(fold_spectral_norm(sd, f"{p}.mlp.fc1"), sd[f"{p}.mlp.fc1.bias"]),
                gemm_weights(fold_spectral_norm(sd, f"{p}.mlp.fc2"), sd[f"{p}.mlp.fc2.bias"]),
                swin_position_bias(sd, f"{p}.attn", geo.ws, geo.num_heads), ls.contiguous(),
                sd[f"{p}.norm1.weight"].float().contiguous(), sd[f"{p}.norm1.bias"].float().contiguous(),
                sd[f"{p}.norm2.weight"].float().contiguous(), sd[f"{p}.norm2.bias"].float().contiguous()))
        up = conv_tc_weights(convt_k2s2_weights(fold_spectral_norm(sd, f"{n}.up.conv", 1), sd[f"{n}.up.conv.bias"]))
        # dense head: output feature (py*pw + px)*C + c (fuxi.py:485); pad C so the GEMM's N is a multiple of 4
        pp, C = geo.patch_height * geo.patch_width, geo.out_chans
        cp = C if (pp * C) % 4 == 0 else _round_up(C, 4)
        wf, bf = fold_spectral_norm(sd, "fc"), sd["fc.bias"].float()
        if cp != C:
            wp = wf.new_zeros((pp, cp, d))
            wp[:, :C] = wf.reshape(pp, C, d)
            bp = bf.new_zeros((pp, cp))
            bp[:, :C] = bf.reshape(pp, C)
            wf, bf = wp.reshape(pp * cp, d), bp.reshape(-1)
        head = gemm_weights(wf, bf)
    return FuxiWeights(cube, sd["cube_embedding.norm.weight"].float().contiguous(),
                       sd["cube_embedding.norm.bias"].float().contiguous(), down, stack(f"{n}.down"), blocks, up,
                       stack(f"{n}.up"), head, cp)


# ---- launch plan ---------------------------------------------------------------------------------------------------

class _FuxiPlan:
    """Workspace + ordered kernel launches of one FuXi forward for a fixed batch size."""

    def __init__(self, geo: FuxiGeometry, wts: FuxiWeights, batch: int, device):
        self.geo, self.batch = geo, batch
        g, B, d = geo, batch, geo.dim
        f32 = dict(device=device, dtype=torch.float32)
        f16 = dict(device=device, dtype=torch.float16)
        self.ld0 = _round_up(g.in_chans * g.frames, 8)
        self.xp = (torch.empty((B, g.h_pad, g.w_pad, self.ld0), **f16), torch.empty((B, g.h_pad, g.w_pad, self.ld0), **f16))
        n_hi = B * g.lat * g.lon * d          # patch-grid fields
        n_lo = B * g.th * g.tw * d            # token-grid fields (before the window padding)
        M = B * g.gh * g.gw                   # tokens of the Swin stage
        self.M = M
        self.e0 = torch.empty(n_hi, **f32)    # cube embedding / UpBlock shortcut u0 (fp32)
        self.a = torch.empty(n_hi, **f32)     # conv outputs awaiting GroupNorm
        self.pa = (torch.empty(n_hi, **f16), torch.empty(n_hi, **f16))   # operand planes on the patch grid (ping)
        self.pb = (torch.empty(n_hi, **f16), torch.empty(n_hi, **f16))   # (pong)
        self.d0 = torch.empty(n_lo, **f32)    # DownBlock conv output (skip of its residual stack)
        self.sc = torch.empty(n_lo, **f32)    # DownBlock output = shortcut of the U (fuxi.py:279)
        self.catp = (torch.empty((B * g.th * g.tw, 2 * d), **f16), torch.empty((B * g.th * g.tw, 2 * d), **f16))
        self.x = torch.empty((M, d), **f32)   # residual stream of the Swin stage (padded token grid)
        self.xpl = (torch.empty((M, d), **f16), torch.empty((M, d), **f16))
        self.t = torch.empty((M, d), **f32)
        self.qkv = torch.empty((M, 3 * d), **f32)
        self.att = (torch.empty((M, d), **f16), torch.empty((M, d), **f16))
        self.hid = (torch.empty((M, 4 * d), **f16), torch.empty((M, 4 * d), **f16))
        self.cp = wts.cp
        self.ytok = torch.empty((B * g.lat * g.lon, g.patch_height * g.patch_width * wts.cp), **f32)
        self.gn_stats = torch.empty((B, g.num_groups, 2), **f32)
        self.gn_scratch = torch.empty(ops.groupnorm_scratch_bytes(B, g.lat * g.lon, d) // 4 + 4, **f32)
        # index lists of the zero pad (token grid -> padded grid) and of the crop (padded grid -> token grid)
        left, right, top, bottom = g.pad2d
        yy, xx = torch.meshgrid(torch.arange(g.gh), torch.arange(g.gw), indexing="ij")
        inside = (yy >= top) & (yy < top + g.th) & (xx >= left) & (xx < left + g.tw)
        src = torch.where(inside, (yy - top) * g.tw + (xx - left), torch.full_like(yy, -1))
        bo = torch.arange(B)[:, None, None]
        self.pad_idx = torch.where(src[None] >= 0, src[None] + bo * (g.th * g.tw), src[None]).reshape(-1).to(
            device=device, dtype=torch.int32)
        ty, tx = torch.meshgrid(torch.arange(g.th), torch.arange(g.tw), indexing="ij")
        crop = ((ty + top) * g.gw + (tx + left))[None] + bo * (g.gh * g.gw)
        self.crop_idx = crop.reshape(-1).to(device=device, dtype=torch.int32)
        self.steps: List[tuple] = []
        self._keep: List[object] = []
        self._build(wts)

    def _add(self, fn, args, tag, flops=0.0, nbytes=0.0):
        self.steps.append((fn, args, tag, float(flops), float(nbytes)))

    def _conv_tc(self, in_hi, in_lo, wts, tag, **kw):
        desc = ops.make_conv_tc_desc(in_hi, in_lo, wts, **kw)
        m = kw["B"] * kw["Ho"] * kw["Wo"]
        self._add(ops.conv_f16x2_tc, (desc,), tag, 2.0 * m * wts.n * wts.t * wts.cin * wts.phases)

    def _gemm(self, a_hi, a_lo, wts, tag, **kw):
        desc = ops.make_gemm_desc(a_hi, a_lo, wts, **kw)
        self._add(ops.gemm_f16x2_tc, (desc,), tag, 2.0 * kw["M"] * wts.n * wts.k)

    def _stack(self, sw: StackWeights, B, h, w, x_f32, xp, tmp_p, out_planes, tag):
        """2 x (conv3x3 + GroupNorm + SiLU) + skip (fuxi.py:160-172, 189-201).  ``x_f32`` / ``xp``: the input as fp32 and as
        planes; result: fp32 into ``out_planes['f32']`` and / or planes."""
        g, d = self.geo, self.geo.dim
        n = B * h * w * d
        self._conv_tc(xp[0], xp[1], sw.convs[0], f"{tag}_conv3x3", B=B, Hi=h, Wi=w, lda=d, Ho=h, Wo=w, out=self.a, ldc=d)
        self._add(ops.groupnorm_silu_f16x2, (self.a, d, self.gn_stats, self.gn_scratch, sw.gn_w[0], sw.gn_b[0], None, 0,
                                             tmp_p[0], tmp_p[1], d, 0, B, h * w, d, g.num_groups), "groupnorm_silu", 0, 8.0 * n)
        self._conv_tc(tmp_p[0], tmp_p[1], sw.convs[1], f"{tag}_conv3x3", B=B, Hi=h, Wi=w, lda=d, Ho=h, Wo=w, out=self.a, ldc=d)
        if out_planes.get("f32") is not None:
            self._add(ops.groupnorm_silu, (self.a, d, self.gn_stats, self.gn_scratch, sw.gn_w[1], sw.gn_b[1], x_f32, d,
                                           out_planes["f32"], d, B, h * w, d, g.num_groups), "groupnorm_silu", 0, 16.0 * n)
        else:
            hi, lo = out_planes["planes"]
            self._add(ops.groupnorm_silu_f16x2, (self.a, d, self.gn_stats, self.gn_scratch, sw.gn_w[1], sw.gn_b[1], x_f32, d,
                                                 hi, lo, d, 0, B, h * w, d, g.num_groups), "groupnorm_silu", 0, 12.0 * n)

    def _build(self, wts: FuxiWeights):
        g, B, d, M = self.geo, self.batch, self.geo.dim, self.M
        add = self._add
        n_hi, n_lo = B * g.lat * g.lon, B * g.th * g.tw
        # CubeEmbedding (fuxi.py:112-143): patch GEMM, LayerNorm over the embedding channel
        self._conv_tc(self.xp[0], self.xp[1], wts.cube, "cube_embed", B=B, Hi=g.h_pad, Wi=g.w_pad, lda=self.ld0, Ho=g.lat,
                      Wo=g.lon, out=self.e0, ldc=d)
        add(ops.layernorm_f16x2, (self.e0, d, self.pa[0], self.pa[1], d, wts.cube_g, wts.cube_b, n_hi, d), "layernorm", 0,
            8.0 * n_hi * d)
        # DownBlock (fuxi.py:146-172): conv3x3 stride 2, then the residual stack on its output
        d0p = (self.pb[0][: n_lo * d], self.pb[1][: n_lo * d])
        tmp = (self.pb[0][n_lo * d: 2 * n_lo * d], self.pb[1][n_lo * d: 2 * n_lo * d])
        self._conv_tc(self.pa[0], self.pa[1], wts.down, "down_conv", B=B, Hi=g.lat, Wi=g.lon, lda=d, Ho=g.th, Wo=g.tw,
                      out=self.d0, ldc=d, out_hi=d0p[0], out_lo=d0p[1], ldh=d)
        self._stack(wts.down_stack, B, g.th, g.tw, self.d0, d0p, tmp, {"f32": self.sc}, "down")
        # shortcut planes = lower half of the concat buffer (fuxi.py:292); zero pad to a window multiple (:281-283)
        add(ops.split_f16x2, (self.sc, d, self.catp[0], self.catp[1], 2 * d, n_lo, d), "split", 0, 8.0 * n_lo * d)
        add(ops.gather_rows_ex, (self.sc, d, self.pad_idx, self.x, d, self.xpl[0], self.xpl[1], d, 0, M, d), "window_pad", 0,
            12.0 * M * d)
        # Swin-V2 stage (timm SwinTransformerV2Stage; oracle/swin_v2.py)
        L = g.ws[0] * g.ws[1]
        for i, bw in enumerate(wts.blocks):
            self._gemm(self.xpl[0], self.xpl[1], bw.qkv, "swin_qkv", M=M, lda=d, out=self.qkv, ldc=3 * d)
            add(ops.swin_window_attention, (self.qkv, 3 * d, bw.bias, bw.logit_scale, self.att[0], self.att[1], None, d, B,
                                            g.gh, g.gw, d, g.num_heads, g.ws, g.block_shift(i)), "swin_attention",
                4.0 * M * L * d, 16.0 * M * d)
            self._gemm(self.att[0], self.att[1], bw.proj, "swin_proj", M=M, lda=d, out=self.t, ldc=d)
            add(ops.layernorm_residual, (self.t, d, self.x, d, self.x, d, self.xpl[0], self.xpl[1], d, bw.n1_g, bw.n1_b, M, d),
                "layernorm_residual", 0, 16.0 * M * d)
            self._gemm(self.xpl[0], self.xpl[1], bw.fc1, "swin_fc1", M=M, lda=d, out_hi=self.hid[0], out_lo=self.hid[1],
                       ldh=4 * d, act=_lib.ACT_GELU)
            self._gemm(self.hid[0], self.hid[1], bw.fc2, "swin_fc2", M=M, lda=4 * d, out=self.t, ldc=d)
            add(ops.layernorm_residual, (self.t, d, self.x, d, self.x, d, self.xpl[0], self.xpl[1], d, bw.n2_g, bw.n2_b, M, d),
                "layernorm_residual", 0, 16.0 * M * d)
        # crop the window padding, concat behind the shortcut (fuxi.py:288-292)
        add(ops.gather_rows_ex, (self.x, d, self.crop_idx, None, 0, self.catp[0], self.catp[1], 2 * d, d, n_lo, d),
            "window_crop", 0, 8.0 * n_lo * d)
        # UpBlock (fuxi.py:175-201): ConvTranspose k2 s2 as 4 output-parity phases, residual stack
        self._conv_tc(self.catp[0], self.catp[1], wts.up, "up_convT", B=B, Hi=g.th, Wi=g.tw, lda=2 * d, Ho=g.th, Wo=g.tw,
                      out=self.e0, ldc=d, out_hi=self.pa[0], out_lo=self.pa[1], ldh=d)
        hp = (torch.empty(n_hi * d, device=self.x.device, dtype=torch.float16),
              torch.empty(n_hi * d, device=self.x.device, dtype=torch.float16))
        self._keep.append(hp)
        self._stack(wts.up_stack, B, g.lat, g.lon, self.e0, self.pa, self.pb, {"planes": hp}, "up")
        # dense head on the channel dimension (fuxi.py:420, 484); un-patchify happens in the output pass
        self._gemm(hp[0], hp[1], wts.head, "head", M=n_hi, lda=d, out=self.ytok, ldc=self.ytok.shape[1])

    def _pad(self, x):
        g = self.geo
        lat, lon, mode = ((g.padding.pad_lat, g.padding.pad_lon, g.padding.mode) if g.padding.activate
                          else ((0, 0), (0, 0), "earth"))
        ops.pad_to_pixel_major_f16x2(x, lat, lon, mode, self.ld0, self.xp[0], self.xp[1])

    def _unpad(self, out):
        g, B = self.geo, self.batch
        pt, pl = (g.padding.pad_lat[0], g.padding.pad_lon[0]) if g.padding.activate else (0, 0)
        ops.unpatchify_unpad_resize_to_nchw(self.ytok, out, B, g.out_chans, self.cp, g.lat, g.lon, g.patch_height,
                                            g.patch_width, pt, pl, g.h_crop, g.w_crop, g.h_out, g.w_out)

    def run(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        g, B = self.geo, self.batch
        self._pad(x)
        for fn, args, _tag, _fl, _by in self.steps:
            fn(*args)
        if out is None:
            out = torch.empty((B, *g.out_shape), device=x.device, dtype=torch.float32)
        self._unpad(out)
        return out

    def run_profiled(self, x: torch.Tensor):
        """One forward with a CUDA-event pair around every launch: [(tag, ms, flops, bytes)] (bench.py roofline)."""
        g, B = self.geo, self.batch
        recs = []

        def timed(tag, flops, nbytes, fn, *args):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(*args)
            e1.record()
            recs.append([tag, (e0, e1), flops, nbytes])

        timed("pad", 0.0, 4.0 * x.numel() + 4.0 * self.xp[0].numel(), self._pad, x)
        for fn, args, tag, fl, by in self.steps:
            timed(tag, fl, by, fn, *args)
        out = torch.empty((B, *g.out_shape), device=x.device, dtype=torch.float32)
        timed("unpad_resize", 0.0, 4.0 * self.ytok.numel() + 4.0 * out.numel(), self._unpad, out)
        torch.cuda.synchronize()
        return out, [(t, ev[0].elapsed_time(ev[1]), fl, by) for t, ev, fl, by in recs]


def fuxi_flops_per_forward(geo: FuxiGeometry) -> Dict[str, float]:
    """2 x MAC of every convolution / matmul of one forward (what torch's FlopCounterMode counts), B = 1."""
    d, M, L = geo.dim, geo.gh * geo.gw, geo.ws[0] * geo.ws[1]
    n_hi, n_lo = geo.lat * geo.lon, geo.th * geo.tw
    fl = {
        "cube_embed": 2.0 * n_hi * d * geo.in_chans * geo.frames * geo.patch_height * geo.patch_width,
        "down": 2.0 * n_lo * d * 9 * d * 3,
        "swin_gemm": geo.depth * 2.0 * M * d * (3 * d + d + 8 * d),
        "swin_attention": geo.depth * 4.0 * M * L * d,
        "up": 2.0 * n_lo * 4 * d * 2 * d + 2.0 * n_hi * d * 9 * d * 2,
        "head": 2.0 * n_hi * d * geo.out_chans * geo.patch_height * geo.patch_width,
    }
    fl["total"] = sum(fl.values())
    return fl


# ---- the module ----------------------------------------------------------------------------------------------------

class FuxiB200(_Base):
    """FuXi forecast step on B200.  Constructor = reference keywords (fuxi.py:321-352)."""

    def __init__(self, init_weights: Optional[bool] = None, **kwargs):
        super().__init__()
        self.geometry = geo = build_fuxi_geometry(**kwargs)
        self.use_interp = geo.interp
        self.use_spectral_norm = geo.use_spectral_norm
        self.use_padding = geo.padding.activate
        self.use_post_block = False
        self.patch_size = (geo.frames, geo.patch_height, geo.patch_width)
        self.input_resolution = (geo.th, geo.tw)
        self.out_chans = geo.out_chans
        self.img_size = (geo.frames, geo.h_pad, geo.w_pad)
        self.img_size_original = (geo.frames, geo.image_height, geo.image_width)
        self.image_height, self.image_width = geo.image_height, geo.image_width
        self.channels, self.surface_channels, self.levels = geo.channels, geo.surface_channels, geo.levels
        if self.use_padding:
            self.padding_opt = PaddingView(geo)
        self._init_seed = int(torch.initial_seed() % (2**31))
        self._lazy_init = not bool(init_weights) if init_weights is not None else True
        init = None if self._lazy_init else synthetic_fuxi_state_dict(geo, self._init_seed)
        for key, (shape, role) in fuxi_state_spec(geo).items():
            parts = key.split(".")
            mod = self
            for p in parts[:-1]:
                if p not in mod._modules:
                    mod.add_module(p, _Holder())
                mod = mod._modules[p]
            val = torch.empty(tuple(shape), dtype=torch.float32) if init is None else init[key]
            if role in ("u", "v"):
                mod.register_buffer(parts[-1], val)
            else:
                mod.register_parameter(parts[-1], nn.Parameter(val, requires_grad=False))
        self.register_load_state_dict_post_hook(FuxiB200._loaded_hook)
        self._prepared: Optional[FuxiWeights] = None
        self._prepared_sig = None
        self._plans: Dict[tuple, _FuxiPlan] = {}
        self._weights_version = 0
        self._domain = None  # set by domain.convert_to_domain_parallel

    @staticmethod
    def _loaded_hook(module, incompatible):
        if not incompatible.missing_keys:
            module._lazy_init = False

    def _materialise(self):
        if not self._lazy_init:
            return
        self._lazy_init = False
        init = synthetic_fuxi_state_dict(self.geometry, self._init_seed)
        with torch.no_grad():
            own = nn.Module.state_dict(self)
            for k, v in init.items():
                own[k].copy_(v)

    def state_dict(self, *args, **kwargs):
        self._materialise()
        return super().state_dict(*args, **kwargs)

    def _signature(self):
        ver = 0
        for t in list(self.parameters()) + list(self.buffers()):
            ver += t._version
        first = next(self.parameters())
        return (ver, first.data_ptr(), str(first.device))

    def refresh_weights(self):
        """Re-fold spectral norm / position bias and re-lay weights (automatic when parameters change)."""
        self._materialise()
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("FuxiB200 parameters must live on a CUDA device (call .cuda()/.to('cuda'))")
        on_host = os.environ.get("WXF_FOLD_DEVICE", "cpu") != "cuda"
        sd = {k: (v.detach().cpu() if on_host else v.detach()) for k, v in self.state_dict().items()}
        prepared = prepare_fuxi(sd, self.geometry)
        self._prepared = _to_device(prepared, dev) if on_host else prepared
        self._prepared_sig = self._signature()
        self._plans.clear()
        self._weights_version += 1

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._prepared = None
        self._plans = {}
        self._weights_version = getattr(self, "_weights_version", 0) + 1
        return out

    def _plan_for(self, x: torch.Tensor):
        if self.training:
            raise NotImplementedError("FuxiB200 implements the eval-mode forecast forward only: call .eval()")
        geo = self.geometry
        if x.dim() != 5 or tuple(x.shape[1:]) != geo.in_shape:
            raise ValueError(f"expected input [B, {', '.join(map(str, geo.in_shape))}], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("FuxiB200 has no CPU path: pass a CUDA tensor")
        if x.dtype != torch.float32:
            x = x.float()
        if self._prepared is None or self._prepared_sig != self._signature():
            self.refresh_weights()
        key = (int(x.shape[0]), x.device.index)
        plan = self._plans.get(key)
        if plan is None:
            with torch.cuda.device(x.device):
                if self._domain is not None:
                    from .fuxi_domain import FuxiDomainPlan

                    if x.shape[0] != 1:
                        raise ValueError("the domain-decomposed forward takes one state at a time (batch 1)")
                    dm = self._domain
                    plan = FuxiDomainPlan(geo, self._prepared, dm.domain_rank, dm.domain_world_size, x.device, dm.domain_group)
                else:
                    plan = _FuxiPlan(geo, self._prepared, int(x.shape[0]), x.device)
                self._plans[key] = plan
        return x.contiguous(), plan

    @torch.no_grad()
    def forward(self, x: torch.Tensor, noise=None, forecast_step=None) -> torch.Tensor:
        x, plan = self._plan_for(x)
        with torch.cuda.device(x.device):
            return plan.run(x)


def fuxi_workload(name: str) -> dict:
    """Constructor kwargs of the FuXi configs (SURVEY.md section 8, rows a17-a19 and BASELINE config #4)."""
    arxiv = dict(frames=2, frame_patch_size=2, levels=16, channels=4, surface_channels=7, input_only_channels=3,
                 output_only_channels=0, patch_height=4, patch_width=4, dim=1024, num_groups=32, num_heads=8, window_size=7,
                 depth=16, use_spectral_norm=True, interp=True, post_conf={"activate": False})
    if name == "fuxi_6h_025deg":
        # config/gen_1/arXiv_2024/fuxi_6h_single_step.yml:98-127 moved to the 721x1440 grid (the reference ships no 0.25 deg
        # FuXi config: SURVEY.md a17): earth padding (40, 39) / (80, 80) -> 800 x 1600 -> 200 x 400 patches ->
        # 100 x 200 tokens -> 105 x 203 with the window padding
        return dict(arxiv, image_height=721, image_width=1440,
                    padding_conf=dict(activate=True, mode="earth", pad_lat=[40, 39], pad_lon=[80, 80]))
    if name == "fuxi_6h_arxiv":  # the config as shipped: 640 x 1280, legacy pad ints -> mirror mode (parser.py:423-430)
        return dict(arxiv, image_height=640, image_width=1280,
                    padding_conf=dict(activate=True, mode="mirror", pad_lat=[80, 80], pad_lon=[80, 80]))
    if name == "fuxi_1deg":  # the same architecture on the 181 x 360 grid (parity test size)
        return dict(arxiv, image_height=181, image_width=360, depth=4,
                    padding_conf=dict(activate=True, mode="earth", pad_lat=[21, 22], pad_lon=[12, 12]))
    raise KeyError(name)
