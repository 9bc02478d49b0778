"""ORACLE — test infrastructure only.  Parity status of THIS file: PINNED AGAINST AN INDEPENDENT IMPLEMENTATION of the same
algorithm (HuggingFace ``transformers`` ``Swinv2Stage``: ``tests/golden/make_golden_swin_hf.py`` -> ``swin_v2_hf.pt``,
``tests/test_swin_v2_vs_hf.py``, bit-identical outputs on randomised parameters incl. shifted blocks and clamped logit
scales); NOT pinned against timm itself, which neither the reference tree nor any image here contains.  What still rests on
reading timm alone: the per-dimension window clamp of ``_calc_window_shift`` (HF clamps with one scalar; FuXi never hits
it: its token grid is padded to window multiples) and the qkv-on-``weight_orig`` call pattern under FuXi's spectral norm.

Restatement of ``timm.models.swin_transformer_v2.SwinTransformerV2Stage`` — the third-party block FuXi's
``UTransformer`` instantiates (``/root/reference/credit/models/fuxi.py:4-5, 250-260``; forward call ``:285-287``).
``timm`` (huggingface/pytorch-image-models) is NOT in the reference tree, is not pinned by the reference
(``pyproject.toml:12-48`` does not list it; it arrives transitively) and is not installed in the build image, so this
file restates the published Swin-V2 algorithm (Liu et al., "Swin Transformer V2", and timm's ``swin_transformer_v2.py``
as of the 0.9 / 1.0 series) from its specification; the real timm module cannot be imported here, HuggingFace's port can:

* windows of ``ws x ws`` tokens, blocks alternate shift 0 / ``ws // 2`` (cyclic roll, ``attn_mask`` = -100 between the
  regions the roll glues together); a window larger than the grid is clamped to the grid and its shift set to 0;
* scaled-cosine attention: ``softmax(normalize(q) normalize(k)^T * exp(min(logit_scale, ln 100)) + 16 sigmoid(bias))``;
* continuous relative position bias: a 2-layer MLP (2 -> 512 -> heads, ReLU) over log-spaced relative coordinates;
* ``qkv`` without bias plus learned ``q_bias`` / ``v_bias`` (the k bias is a zero buffer);
* res-post-norm: ``x = x + norm1(attn(x))``, ``x = x + norm2(mlp(x))``; MLP ratio 4, exact-erf GELU;
* no down-sampling when ``dim == out_dim`` (FuXi's case).

Parameter names follow timm's module tree (``blocks.{i}.attn.{qkv,proj,cpb_mlp.0,cpb_mlp.2,logit_scale,q_bias,v_bias}``,
``blocks.{i}.{norm1,norm2}``, ``blocks.{i}.mlp.{fc1,fc2}``) so that reference checkpoints would load.

One consequence of timm's call pattern under FuXi's spectral-norm hooks is part of the arithmetic: the qkv projection uses
the UN-normalised ``weight_orig`` (see ``_WindowAttention``), every other Linear the hook-normalised weight.

``SwinTransformerV2StageStub`` below is the module-form twin of these functions, the stand-in registered as ``timm`` when
``tests/golden/make_golden_fuxi.py`` imports the UNMODIFIED ``credit/models/fuxi.py``: everything of FuXi that lives in
the reference tree (cube embedding, down/up blocks, head, un-patchify, padding, resize) is thereby pinned against the
reference, while the stage itself stays a restatement on both sides.
"""

from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F
from torch import nn


def to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def clamp_window(resolution: Tuple[int, int], window: Tuple[int, int], shift: Tuple[int, int]):
    """timm ``SwinTransformerV2Block._calc_window_shift``: never a window larger than the grid, no shift then."""
    ws = tuple(r if r <= w else w for r, w in zip(resolution, window))
    ss = tuple(0 if r <= w else s for r, w, s in zip(resolution, ws, shift))
    return ws, ss


def relative_coords_table(ws: Tuple[int, int]) -> torch.Tensor:
    """[1, 2*ws0-1, 2*ws1-1, 2]: sign(c) * log2(|8 c / (ws-1)| + 1) / log2(8) (no pretrained window)."""
    ch = torch.arange(-(ws[0] - 1), ws[0], dtype=torch.float32)
    cw = torch.arange(-(ws[1] - 1), ws[1], dtype=torch.float32)
    table = torch.stack(torch.meshgrid(ch, cw, indexing="ij")).permute(1, 2, 0).contiguous().unsqueeze(0)
    table[..., 0] /= max(ws[0] - 1, 1)
    table[..., 1] /= max(ws[1] - 1, 1)
    table = table * 8
    return torch.sign(table) * torch.log2(table.abs() + 1.0) / math.log2(8)


def relative_position_index(ws: Tuple[int, int]) -> torch.Tensor:
    """[N, N] index of the (dy, dx) between two tokens of a window into the (2ws0-1)*(2ws1-1) table."""
    coords = torch.stack(torch.meshgrid(torch.arange(ws[0]), torch.arange(ws[1]), indexing="ij"))
    flat = coords.flatten(1)
    rel = (flat[:, :, None] - flat[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws[0] - 1
    rel[:, :, 1] += ws[1] - 1
    rel[:, :, 0] *= 2 * ws[1] - 1
    return rel.sum(-1)


def window_partition(x: torch.Tensor, ws: Tuple[int, int]) -> torch.Tensor:
    """[B, H, W, C] -> [B * nW, ws0, ws1, C]."""
    b, h, w, c = x.shape
    x = x.view(b, h // ws[0], ws[0], w // ws[1], ws[1], c)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws[0], ws[1], c)


def window_reverse(win: torch.Tensor, ws: Tuple[int, int], res: Tuple[int, int]) -> torch.Tensor:
    h, w = res
    c = win.shape[-1]
    x = win.view(-1, h // ws[0], w // ws[1], ws[0], ws[1], c)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, h, w, c)


def shift_attn_mask(res: Tuple[int, int], ws: Tuple[int, int], ss: Tuple[int, int]):
    """[nW, N, N] additive mask (0 / -100) of a shifted block, None without shift."""
    if not any(ss):
        return None
    h, w = res
    img = torch.zeros((1, h, w, 1))
    cnt = 0
    for hs in ((0, -ws[0]), (-ws[0], -ss[0]), (-ss[0], None)):
        for wsl in ((0, -ws[1]), (-ws[1], -ss[1]), (-ss[1], None)):
            img[:, hs[0]:hs[1], wsl[0]:wsl[1], :] = cnt
            cnt += 1
    mw = window_partition(img, ws).view(-1, ws[0] * ws[1])
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


def window_attention(xw: torch.Tensor, p: Dict[str, torch.Tensor], heads: int, ws: Tuple[int, int], mask) -> torch.Tensor:
    """timm ``WindowAttention.forward`` (Swin-V2) on [B*nW, N, C]; ``p`` holds the EFFECTIVE (spectral-norm folded)
    weights: qkv_w, q_bias, v_bias, logit_scale, cpb0_w, cpb0_b, cpb2_w, proj_w, proj_b."""
    bw, n, c = xw.shape
    bias = torch.cat((p["q_bias"], torch.zeros_like(p["v_bias"]), p["v_bias"]))
    qkv = F.linear(xw, p["qkv_w"], bias).reshape(bw, n, 3, heads, -1).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    attn = F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1)
    attn = attn * torch.clamp(p["logit_scale"], max=math.log(1.0 / 0.01)).exp()
    table = F.linear(torch.relu(F.linear(relative_coords_table(ws), p["cpb0_w"], p["cpb0_b"])), p["cpb2_w"])
    table = table.view(-1, heads)
    rpb = table[relative_position_index(ws).view(-1)].view(n, n, -1).permute(2, 0, 1).contiguous()
    attn = attn + (16 * torch.sigmoid(rpb)).unsqueeze(0)
    if mask is not None:
        nw = mask.shape[0]
        attn = attn.view(bw // nw, nw, heads, n, n) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, n, n)
    attn = attn.softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(bw, n, c)
    return F.linear(out, p["proj_w"], p["proj_b"])


def block_forward(x: torch.Tensor, p: Dict[str, torch.Tensor], heads: int, ws, ss, mask) -> torch.Tensor:
    """timm ``SwinTransformerV2Block.forward`` on [B, H, W, C] (res-post-norm)."""
    b, h, w, c = x.shape
    sx = torch.roll(x, shifts=(-ss[0], -ss[1]), dims=(1, 2)) if any(ss) else x
    win = window_partition(sx, ws).view(-1, ws[0] * ws[1], c)
    win = window_attention(win, p, heads, ws, mask).view(-1, ws[0], ws[1], c)
    sx = window_reverse(win, ws, (h, w))
    a = torch.roll(sx, shifts=ss, dims=(1, 2)) if any(ss) else sx
    x = x + F.layer_norm(a, (c,), p["norm1_w"], p["norm1_b"], 1e-5)
    y = F.linear(F.gelu(F.linear(x.reshape(b, -1, c), p["fc1_w"], p["fc1_b"])), p["fc2_w"], p["fc2_b"])
    x = x.reshape(b, -1, c) + F.layer_norm(y, (c,), p["norm2_w"], p["norm2_b"], 1e-5)
    return x.reshape(b, h, w, c)


def stage_forward(x: torch.Tensor, blocks, heads: int, resolution: Tuple[int, int], window: int) -> torch.Tensor:
    """``SwinTransformerV2Stage.forward`` with ``dim == out_dim`` (no PatchMerging): the blocks in order, block i shifted
    by ``window // 2`` when i is odd.  ``blocks`` = list of per-block tensor dicts (see ``window_attention``)."""
    win = to_2tuple(window)
    for i, p in enumerate(blocks):
        shift = (0, 0) if i % 2 == 0 else tuple(w // 2 for w in win)
        ws, ss = clamp_window(tuple(resolution), win, shift)
        x = block_forward(x, p, heads, ws, ss, shift_attn_mask(tuple(resolution), ws, ss))
    return x


# ----------------------------------------------------------------------------------------------------------------------
# nn.Module stand-in for the absent timm class (golden generation only)


class _WindowAttention(nn.Module):
    """Module form with timm's CALL PATTERN, which matters under FuXi's old-style spectral-norm hooks (fuxi.py:17-23): the
    hook recomputes ``module.weight`` in a forward PRE-hook, i.e. only when the module itself is called.  timm calls
    ``self.proj(x)``, ``self.cpb_mlp(table)`` and the MLP's layers as modules (hooks fire), but computes the qkv projection
    as ``F.linear(x, self.qkv.weight, cat(q_bias, k_bias, v_bias))`` — the qkv hook never fires and ``qkv.weight`` stays
    the plain attribute the hook registration left behind: ``weight_orig.data``, un-normalised."""

    def __init__(self, dim, heads, ws):
        super().__init__()
        self.heads, self.ws = heads, ws
        self.logit_scale = nn.Parameter(torch.log(10 * torch.ones((heads, 1, 1))))
        self.cpb_mlp = nn.Sequential(nn.Linear(2, 512, bias=True), nn.ReLU(inplace=True), nn.Linear(512, heads, bias=False))
        self.register_buffer("relative_coords_table", relative_coords_table(ws), persistent=False)
        self.register_buffer("relative_position_index", relative_position_index(ws), persistent=False)
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.q_bias = nn.Parameter(torch.zeros(dim))
        self.register_buffer("k_bias", torch.zeros(dim), persistent=False)
        self.v_bias = nn.Parameter(torch.zeros(dim))
        self.proj = nn.Linear(dim, dim)
        self.softmax = nn.Softmax(dim=-1)

    def forward(self, x, mask=None):
        bw, n, c = x.shape
        qkv = F.linear(input=x, weight=self.qkv.weight, bias=torch.cat((self.q_bias, self.k_bias, self.v_bias)))
        q, k, v = qkv.reshape(bw, n, 3, self.heads, -1).permute(2, 0, 3, 1, 4).unbind(0)
        attn = F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1)
        attn = attn * torch.clamp(self.logit_scale, max=math.log(1.0 / 0.01)).exp()
        table = self.cpb_mlp(self.relative_coords_table).view(-1, self.heads)
        rpb = table[self.relative_position_index.view(-1)].view(n, n, -1).permute(2, 0, 1).contiguous()
        attn = attn + (16 * torch.sigmoid(rpb)).unsqueeze(0)
        if mask is not None:
            nw = mask.shape[0]
            attn = (attn.view(bw // nw, nw, self.heads, n, n) + mask.unsqueeze(1).unsqueeze(0)).view(-1, self.heads, n, n)
        attn = self.softmax(attn)
        return self.proj((attn @ v).transpose(1, 2).reshape(bw, n, c))


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _Block(nn.Module):
    def __init__(self, dim, resolution, heads, window, shift, mlp_ratio=4.0):
        super().__init__()
        self.resolution = tuple(resolution)
        self.ws, self.ss = clamp_window(self.resolution, to_2tuple(window), to_2tuple(shift))
        self.attn = _WindowAttention(dim, heads, self.ws)
        self.norm1 = nn.LayerNorm(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))
        self.norm2 = nn.LayerNorm(dim)
        self.register_buffer("attn_mask", shift_attn_mask(self.resolution, self.ws, self.ss), persistent=False)

    def forward(self, x):
        b, h, w, c = x.shape
        sx = torch.roll(x, shifts=(-self.ss[0], -self.ss[1]), dims=(1, 2)) if any(self.ss) else x
        win = window_partition(sx, self.ws).view(-1, self.ws[0] * self.ws[1], c)
        win = self.attn(win, mask=self.attn_mask).view(-1, self.ws[0], self.ws[1], c)
        sx = window_reverse(win, self.ws, self.resolution)
        a = torch.roll(sx, shifts=self.ss, dims=(1, 2)) if any(self.ss) else sx
        x = x + self.norm1(a)
        x = x.reshape(b, -1, c)
        x = x + self.norm2(self.mlp(x))
        return x.reshape(b, h, w, c)


class SwinTransformerV2StageStub(nn.Module):
    """Same constructor positional arguments, parameter tree and module call pattern as timm's stage, as FuXi uses it
    (fuxi.py:250-260)."""

    def __init__(self, dim, out_dim, input_resolution, depth, num_heads, window_size, downsample=False, mlp_ratio=4.0,
                 qkv_bias=True, proj_drop=0.0, attn_drop=0.0, drop_path=0.0, **_):
        super().__init__()
        assert dim == out_dim and not downsample and qkv_bias
        assert proj_drop == 0 and attn_drop == 0 and (drop_path == 0 or drop_path == [0] * depth)
        win = to_2tuple(window_size)
        shift = tuple(w // 2 for w in win)
        self.downsample = nn.Identity()
        self.blocks = nn.ModuleList([_Block(dim, input_resolution, num_heads, win, (0, 0) if i % 2 == 0 else shift, mlp_ratio)
                                     for i in range(depth)])

    def forward(self, x):
        x = self.downsample(x)
        for blk in self.blocks:
            x = blk(x)
        return x


def install_timm_stub():
    """Register the stand-in under the two import paths credit/models/fuxi.py uses (golden generation, build container)."""
    import sys
    import types

    if "timm" in sys.modules and not getattr(sys.modules["timm"], "_wxf_stub", False):
        return False  # a real timm is importable: use it
    timm = types.ModuleType("timm")
    timm._wxf_stub = True
    layers = types.ModuleType("timm.layers")
    helpers = types.ModuleType("timm.layers.helpers")
    helpers.to_2tuple = to_2tuple
    models = types.ModuleType("timm.models")
    sv2 = types.ModuleType("timm.models.swin_transformer_v2")
    sv2.SwinTransformerV2Stage = SwinTransformerV2StageStub
    timm.layers, layers.helpers, timm.models, models.swin_transformer_v2 = layers, helpers, models, sv2
    for name, mod in (("timm", timm), ("timm.layers", layers), ("timm.layers.helpers", helpers), ("timm.models", models),
                      ("timm.models.swin_transformer_v2", sv2)):
        sys.modules[name] = mod
    return True
