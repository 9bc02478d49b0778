"""ORACLE — test infrastructure only.  Parity status: PINNED for everything in the reference tree; the Swin-V2 stage is
pinned against HuggingFace's independent ``Swinv2Stage`` port, not against timm itself (see ``oracle/swin_v2.py``).

A CPU restatement, in plain fp32 PyTorch tensor ops and as pure functions of a *state dict*, of the reference's FuXi
forecast forward step ``y = Fuxi(x)`` (``/root/reference/credit/models/fuxi.py:454-506``): boundary padding, cube
embedding (Conv3d patchify + LayerNorm, ``:82-143``), U-Transformer (DownBlock ``:146-172``, zero pad to a window multiple
``:67-79``, Swin-V2 stage, crop, skip concat, UpBlock ``:175-201``; ``UTransformer.forward`` ``:274-305``), dense head and
un-patchify (``:484-489``), un-pad and bilinear resize (``:491-498``).  ``use_noise`` / ``post_conf`` are off (the noise
branch cannot even be constructed in the reference: SURVEY.md §8c).

The CUDA path (``miles_credit_b200/fuxi.py``) is checked against this file.
Pinning of the in-tree parts: ``tests/golden/make_golden_fuxi.py`` imports the UNMODIFIED ``credit/models/fuxi.py``
through ``credit.models.load_model`` with the Swin-V2 stand-in of ``oracle/swin_v2.py`` registered as ``timm`` and stores
input, state dict and output in ``tests/golden/unit_fuxi.pt``; ``tests/test_fuxi_oracle.py`` checks this file against it.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from oracle import swin_v2
from oracle.crossformer_oracle import bilinear_resize, effective_weight, pad_field, unpad_field


@dataclass
class FuxiSpec:
    """The ``model:`` keywords of the reference constructor that shape the arithmetic (fuxi.py:321-352)."""

    image_height: int
    image_width: int
    patch_height: int
    patch_width: int
    frames: int
    frame_patch_size: int
    levels: int
    channels: int
    surface_channels: int
    input_only_channels: int
    output_only_channels: int
    dim: int
    num_groups: int
    num_heads: int
    depth: int
    window_size: int
    interp: bool = True
    padding: Optional[dict] = None  # {"mode", "pad_lat", "pad_lon"} or None

    @classmethod
    def from_kwargs(cls, **kw):
        pc = kw.get("padding_conf") or {"activate": False}
        padding = None
        if pc.get("activate"):
            padding = dict(mode=pc.get("mode", "earth"), pad_lat=tuple(pc["pad_lat"]), pad_lon=tuple(pc["pad_lon"]))
        names = [f for f in cls.__dataclass_fields__ if f not in ("padding",)]
        return cls(padding=padding, **{n: kw[n] for n in names if n in kw})

    @property
    def in_chans(self) -> int:
        return self.channels * self.levels + self.surface_channels + self.input_only_channels

    @property
    def out_chans(self) -> int:
        return self.channels * self.levels + self.surface_channels + self.output_only_channels

    @property
    def padded_size(self) -> Tuple[int, int]:
        if self.padding is None:
            return self.image_height, self.image_width
        return (self.image_height + sum(self.padding["pad_lat"]), self.image_width + sum(self.padding["pad_lon"]))

    @property
    def input_resolution(self) -> Tuple[int, int]:
        """Token grid of the Swin stage = patches / 2 (fuxi.py:402-405; python ``round``)."""
        h, w = self.padded_size
        return round(h / self.patch_height / 2), round(w / self.patch_width / 2)


def get_pad2d(resolution: Tuple[int, int], window: Tuple[int, int]):
    """(left, right, top, bottom) zero padding to a window multiple (fuxi.py:25-79), smaller half first."""
    lat, lon = resolution
    left = right = top = bottom = 0
    if lat % window[0]:
        p = window[0] - lat % window[0]
        top = p // 2
        bottom = p - top
    if lon % window[1]:
        p = window[1] - lon % window[1]
        left = p // 2
        right = p - left
    return left, right, top, bottom


def cube_embedding(x: torch.Tensor, sd: Dict[str, torch.Tensor], spec: FuxiSpec) -> torch.Tensor:
    """CubeEmbedding.forward (fuxi.py:112-143): Conv3d with kernel = stride = patch (no spectral norm: the hook skips
    Conv3d, :17-23), LayerNorm over the embedding channel, time squeezed (:475)."""
    patch = (spec.frame_patch_size, spec.patch_height, spec.patch_width)
    e = F.conv3d(x, sd["cube_embedding.proj.weight"], sd["cube_embedding.proj.bias"], stride=patch)
    b, d, t, la, lo = e.shape
    e = e.reshape(b, d, -1).transpose(1, 2)
    e = F.layer_norm(e, (d,), sd["cube_embedding.norm.weight"], sd["cube_embedding.norm.bias"], 1e-5)
    return e.transpose(1, 2).reshape(b, d, t, la, lo).squeeze(2)


def residual_stack(x: torch.Tensor, sd, prefix: str, groups: int) -> torch.Tensor:
    """2 x (conv3x3 p1, GroupNorm, SiLU) + skip, shared by DownBlock (:160-172) and UpBlock (:189-201)."""
    y = x
    for ci, gi in ((0, 1), (3, 4)):
        y = F.conv2d(y, effective_weight(sd, f"{prefix}.b.{ci}"), sd[f"{prefix}.b.{ci}.bias"], padding=1)
        y = F.group_norm(y, groups, sd[f"{prefix}.b.{gi}.weight"], sd[f"{prefix}.b.{gi}.bias"], 1e-5)
        y = F.silu(y)
    return y + x


def swin_blocks(sd, prefix: str, depth: int):
    """Per-block tensors of the stage with the spectral-norm hooks of ``apply_spectral_norm`` (fuxi.py:17-23: every
    nn.Linear, including timm's) folded."""
    out = []
    for i in range(depth):
        p = f"{prefix}.blocks.{i}"
        out.append(dict(
            # timm computes F.linear(x, self.qkv.weight, ...) without calling the module, so the spectral-norm pre-hook
            # never fires for qkv: the weight in use is the raw weight_orig (oracle/swin_v2.py, _WindowAttention)
            qkv_w=sd.get(f"{p}.attn.qkv.weight_orig", sd.get(f"{p}.attn.qkv.weight")),
            q_bias=sd[f"{p}.attn.q_bias"], v_bias=sd[f"{p}.attn.v_bias"],
            logit_scale=sd[f"{p}.attn.logit_scale"], cpb0_w=effective_weight(sd, f"{p}.attn.cpb_mlp.0"),
            cpb0_b=sd[f"{p}.attn.cpb_mlp.0.bias"], cpb2_w=effective_weight(sd, f"{p}.attn.cpb_mlp.2"),
            proj_w=effective_weight(sd, f"{p}.attn.proj"), proj_b=sd[f"{p}.attn.proj.bias"],
            norm1_w=sd[f"{p}.norm1.weight"], norm1_b=sd[f"{p}.norm1.bias"],
            fc1_w=effective_weight(sd, f"{p}.mlp.fc1"), fc1_b=sd[f"{p}.mlp.fc1.bias"],
            fc2_w=effective_weight(sd, f"{p}.mlp.fc2"), fc2_b=sd[f"{p}.mlp.fc2.bias"],
            norm2_w=sd[f"{p}.norm2.weight"], norm2_b=sd[f"{p}.norm2.bias"]))
    return out


def u_transformer(x: torch.Tensor, sd, spec: FuxiSpec) -> torch.Tensor:
    """UTransformer.forward (fuxi.py:274-305), ``use_noise`` off."""
    n = "u_transformer"
    x = F.conv2d(x, effective_weight(sd, f"{n}.down.conv"), sd[f"{n}.down.conv.bias"], stride=2, padding=1)
    x = residual_stack(x, sd, f"{n}.down", spec.num_groups)
    shortcut = x
    win = swin_v2.to_2tuple(spec.window_size)
    left, right, top, bottom = get_pad2d(spec.input_resolution, win)
    x = F.pad(x, (left, right, top, bottom))
    res = (spec.input_resolution[0] + top + bottom, spec.input_resolution[1] + left + right)
    tokens = swin_v2.stage_forward(x.permute(0, 2, 3, 1), swin_blocks(sd, f"{n}.layer", spec.depth), spec.num_heads, res,
                                   spec.window_size)
    x = tokens.permute(0, 3, 1, 2)
    x = x[:, :, top: res[0] - bottom, left: res[1] - right]
    x = torch.cat([shortcut, x], dim=1)
    x = F.conv_transpose2d(x, effective_weight(sd, f"{n}.up.conv", sn_dim=1), sd[f"{n}.up.conv.bias"], stride=2)
    return residual_stack(x, sd, f"{n}.up", spec.num_groups)


def forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], spec: FuxiSpec) -> torch.Tensor:
    """Fuxi.forward (fuxi.py:454-506): [B, C_in, T, H, W] -> [B, C_out, 1, H, W]."""
    if spec.padding is not None:
        x = pad_field(x, spec.padding["mode"], spec.padding["pad_lat"], spec.padding["pad_lon"])
    b = x.shape[0]
    lat, lon = (2 * r for r in spec.input_resolution)
    x = cube_embedding(x, sd, spec)
    x = u_transformer(x, sd, spec)
    # dense head on the channel dimension, then un-patchify (:484-489)
    y = F.linear(x.permute(0, 2, 3, 1), effective_weight(sd, "fc"), sd["fc.bias"])
    ph, pw = spec.patch_height, spec.patch_width
    y = y.reshape(b, lat, lon, ph, pw, spec.out_chans).permute(0, 1, 3, 2, 4, 5)
    y = y.reshape(b, lat * ph, lon * pw, spec.out_chans).permute(0, 3, 1, 2)
    if spec.padding is not None:
        y = unpad_field(y, spec.padding["pad_lat"], spec.padding["pad_lon"])
    if spec.interp:
        y = bilinear_resize(y, spec.image_height, spec.image_width)
    return y.unsqueeze(2)
