"""``CrossFormerB200`` — drop-in replacement of ``credit.models.crossformer.CrossFormer`` for the forecast step.

Same constructor keywords (crossformer.py:372-401), same ``state_dict`` keys and shapes (so reference
checkpoints load with ``strict=True``), same tensor contract ``[B, C_in, frames, H, W] -> [B, C_out,
output_frames, H, W]`` (crossformer.py:593-644) and the attributes other CREDIT code reads
(``use_padding``, ``padding_opt``, ``image_height``, ``image_width``, ``use_interp``, ``channels``, ``levels``,
``surface_channels``; trainers/trainer_gen2.py:83-89, base_model.py:44-46).  It subclasses CREDIT's
``BaseModel`` when CREDIT is importable (so ``@register_model`` accepts it, models/__init__.py:154), else a
local mirror of it.

Inside, the forward is a fixed launch plan over hand-written sm_100a kernels reached through the C-ABI
(include/wxformer_b200.h).  Eval-mode forward only: there is no autograd, no CPU path and no PyTorch
fallback — a missing library or a CPU tensor raises.
"""

from __future__ import annotations

import copy
import logging
import os
from typing import Dict, List, Optional

import torch
from torch import nn

from . import lib as _lib
from . import ops
from .geometry import Geometry, build_geometry, check_kernel_limits, state_spec
from .synth import synthetic_state_dict
from .weights import ConvWeights, GemmWeights, PreparedWeights, prepare

logger = logging.getLogger(__name__)

try:  # CREDIT present: be a real BaseModel so credit.models.register_model accepts the class
    from credit.models.base_model import BaseModel as _Base  # type: ignore
except Exception:  # noqa: BLE001 - CREDIT (or one of its heavy deps) absent: local mirror

    class _Base(nn.Module):
        """Mirror of credit/models/base_model.py:12-126 (checkpoint classmethods + reshape helper)."""

        def __init__(self):
            super().__init__()

        def split_and_reshape(self, tensor):
            t1 = tensor[:, : int(self.channels * self.levels), :, :, :]
            t2 = tensor[:, -int(self.surface_channels):, :, :, :]
            t1 = t1.view(t1.shape[0], self.channels, self.levels, t1.shape[2], t1.shape[3], t1.shape[4])
            return t1, t2

        @classmethod
        def load_model(cls, conf):
            conf = copy.deepcopy(conf)
            save_loc = os.path.expandvars(conf["save_loc"])
            ckpt = os.path.join(save_loc, "model_checkpoint.pt")
            if not os.path.isfile(ckpt):
                ckpt = os.path.join(save_loc, "checkpoint.pt")
            return cls._from_checkpoint(conf, ckpt)

        @classmethod
        def load_model_name(cls, conf, model_name):
            conf = copy.deepcopy(conf)
            # base_model.py:94-117: an FSDP checkpoint IS the state dict, every other mode wraps it in "model_state_dict"
            fsdp = conf.get("trainer", {}).get("mode") == "fsdp"
            return cls._from_checkpoint(conf, os.path.join(os.path.expandvars(conf["save_loc"]), model_name),
                                        bare=fsdp, named=True)

        @classmethod
        def _from_checkpoint(cls, conf, ckpt, bare=None, named=False):
            if not os.path.isfile(ckpt):
                raise ValueError((f"No saved checkpoint {ckpt} exists." if named else "No saved checkpoint exists.")
                                 + " You must train a model first. Exiting.")
            checkpoint = torch.load(ckpt, map_location="cpu" if not torch.cuda.is_available() else None)
            conf["model"].pop("type", None)
            model = cls(**conf["model"])
            if bare is None:
                bare = "model_state_dict" not in checkpoint
            sd = checkpoint if bare else checkpoint["model_state_dict"]
            msg = model.load_state_dict(sd, strict=False)
            if msg.unexpected_keys:  # models/checkpoint.py:25-31: unexpected keys raise, missing keys warn
                raise RuntimeError(str(msg))
            if msg.missing_keys:
                logger.warning(f"Loaded partial model {msg}")
            return model

        def save_model(self, conf):
            save_loc = os.path.expandvars(conf["save_loc"])
            torch.save({"model_state_dict": self.state_dict()}, os.path.join(save_loc, "checkpoint.pt"))


class _Holder(nn.Module):
    """Parameter container; exists only so state-dict keys equal the reference's module paths."""


class PaddingView:
    """``model.padding_opt`` as CREDIT code expects it (pad/unpad on [..., H, W]; boundary_padding.py:20-48).

    ``pad`` runs the same CUDA kernel as the forward and returns the reference's NCHW layout.
    """

    def __init__(self, geo: Geometry):
        self.mode = geo.padding.mode
        self.pad_NS = list(geo.padding.pad_lat)
        self.pad_WE = list(geo.padding.pad_lon)

    def pad(self, x: torch.Tensor) -> torch.Tensor:
        shape = x.shape
        x5 = x.reshape(-1, 1, 1, shape[-2], shape[-1]) if x.dim() != 5 else x
        b, c, t = x5.shape[:3]
        pm = ops.pad_to_pixel_major(x5.float(), self.pad_NS, self.pad_WE, self.mode, c * t)
        out = pm.permute(0, 3, 1, 2).reshape(b, c, t, pm.shape[1], pm.shape[2])
        return out.reshape(*shape[:-2], pm.shape[1], pm.shape[2]).contiguous()

    def unpad(self, x: torch.Tensor) -> torch.Tensor:
        h, w = x.shape[-2:]
        return x[..., self.pad_NS[0]: h - self.pad_NS[1], self.pad_WE[0]: w - self.pad_WE[1]]


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def _to_device(obj, dev):
    """Recursively move the tensors of the prepared-weight dataclasses to ``dev`` (plain H2D copies, no kernels)."""
    import dataclasses

    if isinstance(obj, torch.Tensor):
        return obj.to(dev)
    if dataclasses.is_dataclass(obj) and not isinstance(obj, type):
        for f in dataclasses.fields(obj):
            setattr(obj, f.name, _to_device(getattr(obj, f.name), dev))
        return obj
    if isinstance(obj, list):
        return [_to_device(o, dev) for o in obj]
    if isinstance(obj, tuple):
        return tuple(_to_device(o, dev) for o in obj)
    return obj


class _Plan:
    """Workspace + ordered kernel launches of one forward for a fixed batch size.

    ``tensor_cores=True``: contractions run on tcgen05 (f16x2 operand planes); every producer writes the planes its
    consumer needs, the fp32 residual stream stays fp32.  ``False``: every contraction on the exact-fp32 CUDA-core
    kernel (validation path).
    """

    def __init__(self, geo: Geometry, wts: PreparedWeights, batch: int, device, tensor_cores: bool = True, noise=None):
        self.geo, self.batch = geo, batch
        self.noise = noise  # ensemble.NoiseState of CrossFormerWithNoiseB200, or None
        if tensor_cores:
            ok = all(c is not None for brs in wts.embeds_tc[1:] for c in brs) and \
                all(u.up.n % 4 == 0 for u in wts.ups)
            if not ok:
                logger.warning("channel counts not multiples of 4: using the exact-fp32 CUDA-core path")
                tensor_cores = False
        self.tensor_cores = tensor_cores
        if noise is not None and not tensor_cores:
            raise NotImplementedError("the noise-injection variant runs on the tensor-core plan only")
        self.attention_tc = tensor_cores and os.environ.get("WXF_ATTN_TC", "1") != "0"
        self.attn_simt_small = tensor_cores and os.environ.get("WXF_ATTN_SIMT_SMALL", "0") == "1"
        f32 = dict(device=device, dtype=torch.float32)
        f16 = dict(device=device, dtype=torch.float16)
        B = batch
        g = geo
        self.toeplitz = tensor_cores and wts.embed0_toep is not None
        if self.toeplitz:  # stage-0 input as fp16 operand planes, channels padded to 64 (zero) for the 128-byte TMA rows
            self.ld0 = 64
            self.xp = None
            self.xp_planes = (torch.empty((B, g.h_pad, g.w_pad, 64), **f16), torch.empty((B, g.h_pad, g.w_pad, 64), **f16))
        else:
            self.ld0 = wts.cin0_pad
            self.xp = torch.empty((B, g.h_pad, g.w_pad, self.ld0), **f32)
        big = max(B * s.h * s.w * s.dim for s in g.stages)
        big = max(big, max(4 * B * u.h_in * u.w_in * u.c_out for u in g.ups))
        big = _round_up(big, 8)
        self.ln = torch.empty(big, **f32)
        y_elems = B * g.h_dec * g.w_dec * g.output_channels
        self.scratch = torch.empty(_round_up(max(4 * big, 2 * big + y_elems, 2 * y_elems + 8), 8), **f32)
        self.ln16 = self.ln.view(torch.float16)            # [hi plane | lo plane], each ln.numel() halves
        self.scratch16 = self.scratch.view(torch.float16)  # [hi plane | lo plane], each scratch.numel() halves
        # residual streams: stages 0..2 live in the upper half of their skip-concat buffer
        self.cat = [torch.empty((B, s.h, s.w, 2 * s.dim), **f32) for s in g.stages[:3]]
        self.x3 = torch.empty((B, g.stages[3].h, g.stages[3].w, g.stages[3].dim), **f32)
        if tensor_cores:
            self.catp = [(torch.empty((B, s.h, s.w, 2 * s.dim), **f16), torch.empty((B, s.h, s.w, 2 * s.dim), **f16))
                         for s in g.stages[:3]]
            s3 = g.stages[3]
            self.x3p = (torch.empty((B, s3.h, s3.w, s3.dim), **f16), torch.empty((B, s3.h, s3.w, s3.dim), **f16))
            nmax = max(4 * B * u.h_in * u.w_in * u.c_out for u in g.ups)
            self.shortp = (torch.empty(nmax, **f16), torch.empty(nmax, **f16))
        self.gn_stats = torch.empty((B, g.dim[0], 2), **f32)
        gn_bytes = max(ops.groupnorm_scratch_bytes(B, 4 * u.h_in * u.w_in, u.c_out) for u in g.ups)
        self.gn_scratch = torch.empty(gn_bytes // 4 + 4, **f32)
        self.steps: List[tuple] = []
        self.bias_tiles: List[torch.Tensor] = []
        self._keep: List[object] = []  # weight views referenced by raw pointer from the launch descriptors
        self._build(wts)

    # -- helpers -------------------------------------------------------------------------------
    def _add(self, fn, args, tag, flops=0.0, nbytes=0.0):
        self.steps.append((fn, args, tag, float(flops), float(nbytes)))

    def _conv(self, inp, wts: ConvWeights, out, tag="conv", **kw):
        desc = ops.make_conv_desc(inp, wts, out, **kw)
        m = kw["B"] * kw["Ho"] * kw["Wo"]
        flops = 2.0 * m * wts.n * wts.t * wts.cin * wts.phases
        self._add(ops.conv_igemm_f32, (desc,), tag, flops)

    def _gemm(self, a_hi, a_lo, wts, tag, **kw):
        desc = ops.make_gemm_desc(a_hi, a_lo, wts, **kw)
        self._add(ops.gemm_f16x2_tc, (desc,), tag, 2.0 * kw["M"] * wts.n * wts.k)

    def _conv_tc(self, in_hi, in_lo, wts, tag, **kw):
        desc = ops.make_conv_tc_desc(in_hi, in_lo, wts, **kw)
        m = kw["B"] * kw["Ho"] * kw["Wo"]
        self._add(ops.conv_f16x2_tc, (desc,), tag, 2.0 * m * wts.n * wts.t * wts.cin * wts.phases)

    def _noise(self, site, x, ldx, out, ldo, planes, B, HW, C):
        """One injection site of the ensemble variant (crossformer_ensemble.py:139-162): the per-(batch, channel)
        coefficient, then feature + eps * coef as fp32 (``out``) and / or operand planes (``planes`` = (hi, lo, ld))."""
        ns = self.noise
        hi, lo, ldh = planes if planes is not None else (None, None, 0)
        self._add(ns.coef, (site, B), "noise_coef", 0, 0)
        self._add(ns.inject, (site, x, ldx, out, ldo, hi, lo, ldh, B, HW, C), "noise_inject", 0,
                  (8.0 if out is not None else 4.0) * B * HW * C + (4.0 * B * HW * C if hi is not None else 0.0))

    def _transformer(self, layers, st, B, h, w, xv, ld, out_planes):
        """Launches of one Transformer stage (crossformer.py:358-365) on the residual stream ``xv`` ([B, h, w, d] view with
        pixel stride ``ld``), in place.  ``out_planes`` = (hi, lo, stride): also emit the stage output as operand planes."""
        g = self.geo
        add = self._add
        tc = self.tensor_cores
        s, d = st.index, st.dim
        m = B * h * w
        scale = float(g.dim_head) ** -0.5
        if out_planes is not None:
            xp_hi, xp_lo, pld = out_planes
        wts_blocks = layers
        ln = self.ln[: m * d]
        wide = self.scratch[: m * 4 * d]
        # fp16 operand planes alias the same workspaces (2 planes x 2 bytes = the fp32 footprint)
        ln_hi, ln_lo = self.ln16[: m * d], self.ln16[self.ln.numel(): self.ln.numel() + m * d]
        hid_off = self.scratch.numel()
        hid_hi, hid_lo = self.scratch16[: m * 4 * d], self.scratch16[hid_off: hid_off + m * 4 * d]
        n_layers = len(wts_blocks)
        for li, layer in enumerate(wts_blocks):
            for half, (att, ff) in enumerate(((layer[0], layer[1]), (layer[2], layer[3]))):
                L = att.wsz * att.wsz
                attn_cost = (4.0 * m * L * d, 16.0 * m * d)
                last = li == n_layers - 1 and half == 1
                if tc:
                    add(ops.layernorm_f16x2, (xv, ld, ln_hi, ln_lo, d, att.ln_g, att.ln_b, m, d), f"layernorm.s{s}", 0,
                        8.0 * m * d)
                    if L == 1:
                        # a one-token window: softmax over a single score is exactly 1, so the attention output is v
                        # (crossformer.py:275-295 with i = j = 1).  Only the v rows of to_qkv are computed and the
                        # out-projection reads them directly: no q, k, and no attention launch.
                        vw = GemmWeights(att.qkv_tc.w_hi[2 * d: 3 * d], att.qkv_tc.w_lo[2 * d: 3 * d], None, d,
                                         att.qkv_tc.k, att.qkv_tc.scale_log2)
                        self._keep.append(vw)
                        v_hi, v_lo = self.scratch16[: m * d], self.scratch16[hid_off: hid_off + m * d]
                        self._gemm(ln_hi, ln_lo, vw, f"qkv.s{s}", M=m, lda=d, out_hi=v_hi, out_lo=v_lo, ldh=d)
                        self._gemm(v_hi, v_lo, att.out_tc, f"out_proj.s{s}", M=m, lda=d, out=xv, ldc=ld, res=xv, ldr=ld)
                    elif self.attn_simt_small and att.kind == _lib.ATTN_LONG and L <= 8:
                        # round-2 candidate (WXF_ATTN_SIMT_SMALL=1): dilated groups of <= 8 tokens (stage 2: L = 4) fill 3 %
                        # of a 128x128 tensor-core tile; the CUDA-core kernel (one thread per query) only has to stream
                        # the tensor.  Both kernels are validated; which is faster at L = 4 is not measured yet.
                        self._gemm(ln_hi, ln_lo, att.qkv_tc, f"qkv.s{s}", M=m, lda=d, out=wide, ldc=3 * d)
                        add(ops.window_attention_f16x2, (wide, 3 * d, att.bias_t, ln_hi, ln_lo, d, B, h, w, d,
                                                         g.dim_head, att.wsz, att.kind, scale), f"attention.s{s}",
                            *attn_cost)
                    elif self.attention_tc and L <= 128:
                        q_hi, q_lo = self.scratch16[: m * 3 * d], self.scratch16[hid_off: hid_off + m * 3 * d]
                        self._gemm(ln_hi, ln_lo, att.qkv_tc, f"qkv.s{s}", M=m, lda=d, out_hi=q_hi, out_lo=q_lo, ldh=3 * d)
                        tile = ops.attention_bias_tile(att.bias_t, w, att.wsz, att.kind)
                        self.bias_tiles.append(tile)
                        add(ops.window_attention_tc, (q_hi, q_lo, 3 * d, tile, ln_hi, ln_lo, d, B, h, w,
                                                      d, g.dim_head, att.wsz, att.kind, scale), f"attention.s{s}",
                            *attn_cost)
                    else:
                        self._gemm(ln_hi, ln_lo, att.qkv_tc, f"qkv.s{s}", M=m, lda=d, out=wide, ldc=3 * d)
                        add(ops.window_attention_f16x2, (wide, 3 * d, att.bias_t, ln_hi, ln_lo, d, B, h, w, d,
                                                         g.dim_head, att.wsz, att.kind, scale), f"attention.s{s}",
                            *attn_cost)
                    if L != 1:
                        self._gemm(ln_hi, ln_lo, att.out_tc, f"out_proj.s{s}", M=m, lda=d, out=xv, ldc=ld, res=xv, ldr=ld)
                    add(ops.layernorm_f16x2, (xv, ld, ln_hi, ln_lo, d, ff.ln_g, ff.ln_b, m, d), f"layernorm.s{s}", 0,
                        8.0 * m * d)
                    self._gemm(ln_hi, ln_lo, ff.fc1_tc, f"ff1.s{s}", M=m, lda=d, out_hi=hid_hi, out_lo=hid_lo, ldh=4 * d,
                               act=_lib.ACT_GELU)
                    if last and out_planes is not None:  # the stage output also feeds the next cross-embed / the decoder
                        self._gemm(hid_hi, hid_lo, ff.fc2_tc, f"ff2.s{s}", M=m, lda=4 * d, out=xv, ldc=ld, res=xv, ldr=ld,
                                   out_hi=xp_hi, out_lo=xp_lo, ldh=pld)
                    else:
                        self._gemm(hid_hi, hid_lo, ff.fc2_tc, f"ff2.s{s}", M=m, lda=4 * d, out=xv, ldc=ld, res=xv, ldr=ld)
                    continue
                add(ops.layernorm, (xv, ld, ln, d, att.ln_g, att.ln_b, m, d), f"layernorm.s{s}", 0, 8.0 * m * d)
                self._conv(ln, att.qkv, wide, tag=f"qkv.s{s}", B=B, Hi=h, Wi=w, lda=d, Ho=h, Wo=w, ldc=3 * d)
                add(ops.window_attention_f32, (wide, 3 * d, att.bias_t, ln, d, B, h, w, d, g.dim_head,
                                               att.wsz, att.kind, scale), f"attention.s{s}", *attn_cost)
                self._conv(ln, att.out, xv, tag=f"out_proj.s{s}", B=B, Hi=h, Wi=w, lda=d, Ho=h, Wo=w, ldc=ld,
                           res=xv, ldr=ld)
                add(ops.layernorm, (xv, ld, ln, d, ff.ln_g, ff.ln_b, m, d), f"layernorm.s{s}", 0, 8.0 * m * d)
                self._conv(ln, ff.fc1, wide, tag=f"ff1.s{s}", B=B, Hi=h, Wi=w, lda=d, Ho=h, Wo=w, ldc=4 * d,
                           act=_lib.ACT_GELU)
                self._conv(wide, ff.fc2, xv, tag=f"ff2.s{s}", B=B, Hi=h, Wi=w, lda=4 * d, Ho=h, Wo=w, ldc=ld,
                           res=xv, ldr=ld)

    def _build(self, wts: PreparedWeights):
        g, B = self.geo, self.batch
        add = self._add
        tc = self.tensor_cores
        src, src_ld, src_h, src_w = self.xp, self.ld0, g.h_pad, g.w_pad
        src_planes = None
        scale = float(g.dim_head) ** -0.5
        for st in g.stages:
            s, d = st.index, st.dim
            if s < 3:
                xbuf, ld, xoff = self.cat[s], 2 * d, d
            else:
                xbuf, ld, xoff = self.x3, d, 0
            xv = xbuf[..., xoff:]  # view: data_ptr carries the channel offset
            m = B * st.h * st.w
            if tc:
                if s < 3:
                    xp_hi, xp_lo, pld = self.catp[s][0][..., d:], self.catp[s][1][..., d:], 2 * d
                else:
                    xp_hi, xp_lo, pld = self.x3p[0], self.x3p[1], d
            for bi, (br, bw) in enumerate(zip(st.branches, wts.embeds[s])):
                if s == 0 and self.toeplitz:
                    tw = wts.embed0_toep[bi]
                    desc = ops.make_toeplitz_desc(self.xp_planes[0], self.xp_planes[1], tw, xbuf, B=B, Hi=src_h, Wi=src_w,
                                                  lda=64, Ho=st.h, Wo=st.w, ldc=ld, c_off=xoff + br.c_off)
                    self._add(ops.cross_embed_toeplitz_tc, (desc,), f"embed0.k{br.kernel}",
                              2.0 * m * br.c_out * st.c_in * br.kernel * br.kernel)
                elif tc and s > 0:
                    self._conv_tc(src_planes[0], src_planes[1], wts.embeds_tc[s][bi], f"embed{s}.k{br.kernel}", B=B,
                                  Hi=src_h, Wi=src_w, lda=src_ld, Ho=st.h, Wo=st.w, out=xbuf, ldc=ld,
                                  c_off=xoff + br.c_off)
                else:
                    self._conv(src, bw, xbuf, tag=f"embed{s}.k{br.kernel}", B=B, Hi=src_h, Wi=src_w, lda=src_ld,
                               Ho=st.h, Wo=st.w, ldc=ld, c_off=xoff + br.c_off)
            self._transformer(wts.blocks[s], st, B, st.h, st.w, xv, ld, (xp_hi, xp_lo, pld) if tc else None)
            if self.noise is not None and self.noise.encoder and s < 3:
                # encoder noise: the stage output (skip connection AND next cross-embed input) is perturbed in place
                self._noise(s, xv, ld, xv, ld, (xp_hi, xp_lo, pld), B, st.h * st.w, d)
            src, src_ld, src_h, src_w = xv, ld, st.h, st.w
            if tc:
                src_planes = (xp_hi, xp_lo)

        # decoder: UpBlock x3 (crossformer.py:107-122) or UpBlockPS x3 (wxformer/crossformer.py:137-162);
        # outputs land in the lower half of the skip buffers
        wx = g.variant == "wxformer"
        st0 = g.stages[0]
        big2 = 2 * (B * st0.h * st0.w * st0.dim)
        y_elems = B * g.h_dec * g.w_dec * g.output_channels
        if wx:
            # up_block4 of the wxformer variant stages its PixelShuffle output (y_elems values, fp32 or hi/lo planes) at
            # scratch[0:y_elems] and the next conv3x3 reads it with a halo while writing y_dec: keep y_dec behind it
            # (output_channels > dim[0]/2 in every shipped wxformer config)
            big2 = max(big2, _round_up(y_elems, 8))
        self.y_dec = self.scratch[big2: big2 + y_elems]
        head_tc = tc and wts.head_tc is not None and (not wx or wts.head2_tc is not None)
        dec_in, dec_ld = self.x3, g.stages[3].dim
        dec_planes = self.x3p if tc else None
        for up, uw, skip in zip(g.ups, wts.ups, (2, 1, 0)):
            ho, wo, c = 2 * up.h_in, 2 * up.w_in, up.c_out
            n = B * ho * wo * c
            short = self.ln[:n]
            a, b = self.scratch[:n], self.scratch[n: 2 * n]
            dst = self.cat[skip]
            if tc:
                sp_hi, sp_lo = self.shortp[0][:n], self.shortp[1][:n]
                b_hi, b_lo = self.scratch16[2 * n: 3 * n], self.scratch16[3 * n: 4 * n]
                if wx:
                    # sub-pixel conv + PixelShuffle -> u (fp32 in the `b` area + planes in the `a` area), then
                    # x = u + sharp(u) -> shortcut (fp32 + planes)
                    u_hi, u_lo = self.scratch16[:n], self.scratch16[n: 2 * n]
                    self._conv_tc(dec_planes[0], dec_planes[1], uw.up_tc, "dec_up", B=B, Hi=up.h_in, Wi=up.w_in,
                                  lda=dec_ld, Ho=up.h_in, Wo=up.w_in, out=b, ldc=c, out_hi=u_hi, out_lo=u_lo, ldh=c)
                    self._conv_tc(u_hi, u_lo, uw.sharp_tc, "dec_conv3x3", B=B, Hi=ho, Wi=wo, lda=c, Ho=ho, Wo=wo,
                                  out=short, ldc=c, res=b, ldr=c, out_hi=sp_hi, out_lo=sp_lo, ldh=c)
                else:
                    self._conv_tc(dec_planes[0], dec_planes[1], uw.up_tc, "dec_up", B=B, Hi=up.h_in, Wi=up.w_in,
                                  lda=dec_ld, Ho=up.h_in, Wo=up.w_in, out=short, ldc=c, out_hi=sp_hi, out_lo=sp_lo, ldh=c)
                self._conv_tc(sp_hi, sp_lo, uw.convs_tc[0], "dec_conv3x3", B=B, Hi=ho, Wi=wo, lda=c, Ho=ho, Wo=wo, out=a,
                              ldc=c)
                add(ops.groupnorm_silu_f16x2, (a, c, self.gn_stats, self.gn_scratch, uw.gn_w[0], uw.gn_b[0], None, 0,
                                               b_hi, b_lo, c, 0, B, ho * wo, c, up.groups), "groupnorm_silu", 0, 8.0 * n)
                self._conv_tc(b_hi, b_lo, uw.convs_tc[1], "dec_conv3x3", B=B, Hi=ho, Wi=wo, lda=c, Ho=ho, Wo=wo, out=a,
                              ldc=c)
                if self.noise is not None:
                    # decoder noise (noise_inject1..3): UpBlock output -> fp32 temporary -> + eps * coef -> concat buffer
                    add(ops.groupnorm_silu, (a, c, self.gn_stats, self.gn_scratch, uw.gn_w[1], uw.gn_b[1], short, c, b, c, B,
                                             ho * wo, c, up.groups), "groupnorm_silu", 0, 16.0 * n)
                    if skip == 0 and not head_tc:
                        self._noise(5 - skip, b, c, dst, 2 * c, None, B, ho * wo, c)
                    else:
                        self._noise(5 - skip, b, c, None, 0, (self.catp[skip][0], self.catp[skip][1], 2 * c), B, ho * wo, c)
                elif skip == 0 and not head_tc:  # fp32 head (odd channel count): keep the fp32 concat buffer
                    add(ops.groupnorm_silu, (a, c, self.gn_stats, self.gn_scratch, uw.gn_w[1], uw.gn_b[1], short, c,
                                             dst, 2 * c, B, ho * wo, c, up.groups), "groupnorm_silu", 0, 16.0 * n)
                else:
                    add(ops.groupnorm_silu_f16x2, (a, c, self.gn_stats, self.gn_scratch, uw.gn_w[1], uw.gn_b[1], short,
                                                   c, self.catp[skip][0], self.catp[skip][1], 2 * c, 0, B, ho * wo, c,
                                                   up.groups), "groupnorm_silu", 0, 12.0 * n)
                dec_planes, dec_ld = self.catp[skip], 2 * c
                dec_in = dst
                continue
            if wx:
                self._conv(dec_in, uw.up, b, tag="dec_up", B=B, Hi=up.h_in, Wi=up.w_in, lda=dec_ld, Ho=up.h_in,
                           Wo=up.w_in, ldc=c)
                self._conv(b, uw.sharp, short, tag="dec_conv3x3", B=B, Hi=ho, Wi=wo, lda=c, Ho=ho, Wo=wo, ldc=c, res=b,
                           ldr=c)
            else:
                self._conv(dec_in, uw.up, short, tag="dec_up", B=B, Hi=up.h_in, Wi=up.w_in, lda=dec_ld, Ho=up.h_in,
                           Wo=up.w_in, ldc=c)
            self._conv(short, uw.convs[0], a, tag="dec_conv3x3", B=B, Hi=ho, Wi=wo, lda=c, Ho=ho, Wo=wo, ldc=c)
            add(ops.groupnorm_silu, (a, c, self.gn_stats, self.gn_scratch, uw.gn_w[0], uw.gn_b[0], None, 0, b, c, B,
                                     ho * wo, c, up.groups), "groupnorm_silu", 0, 12.0 * n)
            self._conv(b, uw.convs[1], a, tag="dec_conv3x3", B=B, Hi=ho, Wi=wo, lda=c, Ho=ho, Wo=wo, ldc=c)
            add(ops.groupnorm_silu, (a, c, self.gn_stats, self.gn_scratch, uw.gn_w[1], uw.gn_b[1], short, c, dst,
                                     2 * c, B, ho * wo, c, up.groups), "groupnorm_silu", 0, 16.0 * n)
            dec_in, dec_ld = dst, 2 * c
        co = g.output_channels
        nv = B * g.h_dec * g.w_dec * co  # up_block4's PixelShuffle output (wxformer variant)
        if wx and head_tc:
            v_hi, v_lo = self.scratch16[:nv], self.scratch16[nv: 2 * nv]
            self._conv_tc(dec_planes[0], dec_planes[1], wts.head_tc, "dec_head", B=B, Hi=st0.h, Wi=st0.w, lda=dec_ld,
                          Ho=st0.h, Wo=st0.w, out_hi=v_hi, out_lo=v_lo, ldh=co)
            self._conv_tc(v_hi, v_lo, wts.head2_tc, "dec_head", B=B, Hi=g.h_dec, Wi=g.w_dec, lda=co, Ho=g.h_dec,
                          Wo=g.w_dec, out=self.y_dec, ldc=co)
        elif wx:
            v = self.scratch[:nv]
            self._conv(dec_in, wts.head, v, tag="dec_head", B=B, Hi=st0.h, Wi=st0.w, lda=dec_ld, Ho=st0.h, Wo=st0.w,
                       ldc=co)
            self._conv(v, wts.head2, self.y_dec, tag="dec_head", B=B, Hi=g.h_dec, Wi=g.w_dec, lda=co, Ho=g.h_dec,
                       Wo=g.w_dec, ldc=co)
        elif head_tc:
            self._conv_tc(dec_planes[0], dec_planes[1], wts.head_tc, "dec_head", B=B, Hi=st0.h, Wi=st0.w, lda=dec_ld,
                          Ho=st0.h, Wo=st0.w, out=self.y_dec, ldc=g.output_channels)
        else:
            self._conv(dec_in, wts.head, self.y_dec, tag="dec_head", B=B, Hi=st0.h, Wi=st0.w, lda=dec_ld, Ho=st0.h,
                       Wo=st0.w, ldc=g.output_channels)

    def _pad(self, x):
        g = self.geo
        lat, lon, mode = ((g.padding.pad_lat, g.padding.pad_lon, g.padding.mode) if g.padding.activate
                          else ((0, 0), (0, 0), "earth"))
        if self.toeplitz:
            ops.pad_to_pixel_major_f16x2(x, lat, lon, mode, 64, self.xp_planes[0], self.xp_planes[1])
        else:
            ops.pad_to_pixel_major(x, lat, lon, mode, self.ld0, out=self.xp)

    def _pad_fields(self, tab):
        """The pre-blocks fused into the padding pass (pipeline.FusedPreblocks.tables): per-variable planes in, z-scored,
        concatenated and padded operand planes out — the [B, C, T, H, W] input tensor never exists."""
        g = self.geo
        table, mean, std, B, C, T, H, W, _keep = tab
        if (B, C * T, T, H, W) != (self.batch, g.input_channels, g.frames, g.image_height, g.image_width):
            raise ValueError(f"fields [B={B}, C={C}, T={T}, {H}, {W}] do not match the model input {(self.batch, *g.in_shape)}")
        lat, lon, mode = ((g.padding.pad_lat, g.padding.pad_lon, g.padding.mode) if g.padding.activate
                          else ((0, 0), (0, 0), "earth"))
        if self.toeplitz:
            ops.preblock_pad_to_pixel_major(table, mean, std, B, C, T, H, W, lat, lon, mode, 64, out_hi=self.xp_planes[0],
                                            out_lo=self.xp_planes[1])
        elif self.ld0 % 8 == 0:
            ops.preblock_pad_to_pixel_major(table, mean, std, B, C, T, H, W, lat, lon, mode, self.ld0, out=self.xp)
        else:
            raise NotImplementedError("the fused pre-blocks need a padded-input channel stride that is a multiple of 8")

    def _unpad(self, out, post=None):
        g, B = self.geo, self.batch
        pt, pl = (g.padding.pad_lat[0], g.padding.pad_lon[0]) if g.padding.activate else (0, 0)
        if post is not None:  # inverse scaling + tracer clamps in the epilogue (pipeline.FusedPostblocks)
            ops.unpad_resize_post_to_nchw(self.y_dec, g.output_channels, out, B, g.output_channels, g.h_dec, g.w_dec, pt, pl,
                                          g.h_crop, g.w_crop, g.h_out, g.w_out, post.scale, post.shift, post.lo, post.hi)
            return
        ops.unpad_resize_to_nchw(self.y_dec, g.output_channels, out, B, g.output_channels, g.h_dec, g.w_dec, pt, pl,
                                 g.h_crop, g.w_crop, g.h_out, g.w_out)

    def run(self, x: Optional[torch.Tensor], out: Optional[torch.Tensor] = None, fields=None, post=None) -> torch.Tensor:
        g, B = self.geo, self.batch
        if fields is not None:
            self._pad_fields(fields)
        else:
            self._pad(x)
        for fn, args, _tag, _fl, _by in self.steps:
            fn(*args)
        if out is None:
            dev = x.device if x is not None else fields[0].device
            out = torch.empty((B, g.base_output_channels, g.output_frames, g.h_out, g.w_out), device=dev, dtype=torch.float32)
        if post is not None:
            self._unpad(out, post)
        else:
            self._unpad(out)
        if self.noise is not None:
            self.noise.advance()
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

        n_in = float(x.numel()) * 4
        n_pad = 4.0 * (self.xp.numel() if self.xp is not None else self.xp_planes[0].numel())
        timed("pad", 0.0, n_in + n_pad, self._pad, x)
        for fn, args, tag, fl, by in self.steps:
            timed(tag, fl, by, fn, *args)
        out = torch.empty((B, g.base_output_channels, g.output_frames, g.h_out, g.w_out), device=x.device,
                          dtype=torch.float32)
        timed("unpad_resize", 0.0, 8.0 * out.numel(), self._unpad, out)
        torch.cuda.synchronize()
        return out, [(t, ev[0].elapsed_time(ev[1]), fl, by) for t, ev, fl, by in recs]


class CrossFormerB200(_Base):
    """WXFormer/CrossFormer forecast step on B200.  Constructor = reference keywords (crossformer.py:372-401)."""

    VARIANT = "crossformer"

    def __init__(self, init_weights: Optional[bool] = None, **kwargs):
        super().__init__()
        kwargs.setdefault("variant", self.VARIANT)
        self.geometry = geo = build_geometry(**kwargs)
        for note in check_kernel_limits(geo):
            logger.warning(note)
        # attributes CREDIT reads off the model
        self.image_height, self.image_width = geo.image_height, geo.image_width
        self.patch_height = self.patch_width = 1
        self.frames, self.output_frames = geo.frames, geo.output_frames
        self.channels, self.levels, self.surface_channels = geo.channels, geo.levels, geo.surface_channels
        self.input_only_channels = geo.input_only_channels
        self.base_input_channels, self.input_channels = geo.base_input_channels, geo.input_channels
        self.base_output_channels, self.output_channels = geo.base_output_channels, geo.output_channels
        self.use_spectral_norm = geo.use_spectral_norm
        self.use_interp = geo.interp
        self.use_padding = geo.padding.activate
        self.use_post_block = False
        self.upsample_v_conv = False
        if self.use_padding:
            self.padding_opt = PaddingView(geo)
        # parameters: same dotted names as the reference module tree.  The reference's constructor leaves a randomly
        # initialised model; here that initialisation (synthetic weights with power-iterated spectral-norm vectors, 124 M
        # parameters at 0.25 deg) is LAZY: it runs at the first forward / state_dict() unless a checkpoint has been loaded by
        # then, which is what BaseModel.load_model does right after construction (base_model.py:72-85).
        self._init_seed = int(torch.initial_seed() % (2**31))
        self._lazy_init = not bool(init_weights) if init_weights is not None else True
        init = None if self._lazy_init else synthetic_state_dict(geo, seed=self._init_seed, sn_iters=5)
        for key, (shape, role) in state_spec(geo).items():
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
        self.register_load_state_dict_post_hook(CrossFormerB200._loaded_hook)
        # exact-fp32 CUDA-core contractions instead of the f16x2 tensor-core GEMMs (validation aid, ~5x slower)
        self.exact_fp32 = os.environ.get("WXF_EXACT_FP32", "0") == "1"
        self._prepared: Optional[PreparedWeights] = None
        self._prepared_sig = None
        self._plans: Dict[tuple, _Plan] = {}
        self._domain = None  # set by domain.convert_to_domain_parallel

    # -- lazy initialisation ---------------------------------------------------------------------------
    @staticmethod
    def _loaded_hook(module, incompatible):
        if not incompatible.missing_keys:
            module._lazy_init = False

    def _materialise(self):
        """Give a model that never received a checkpoint its synthetic initial weights (what the eager constructor did)."""
        if not self._lazy_init:
            return
        self._lazy_init = False
        init = synthetic_state_dict(self.geometry, seed=self._init_seed, sn_iters=5)
        with torch.no_grad():
            own = nn.Module.state_dict(self)
            for k, v in init.items():
                own[k].copy_(v)

    def state_dict(self, *args, **kwargs):
        self._materialise()
        return super().state_dict(*args, **kwargs)

    # -- weight folding ------------------------------------------------------------------------------
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
            raise RuntimeError("CrossFormerB200 parameters must live on a CUDA device (call .cuda()/.to('cuda'))")
        # The fold (spectral norm, position-bias MLPs, plane split) runs ONCE per load, on the host: a few seconds of CPU
        # work instead of ~1000 small ATen launches, the result is bit-identical on every rank, and the only kernels this
        # module ever launches on the GPU are its own (WXF_FOLD_DEVICE=cuda folds on the device instead).
        on_host = os.environ.get("WXF_FOLD_DEVICE", "cpu") != "cuda"
        sd = {k: (v.detach().cpu() if on_host else v.detach()) for k, v in self.state_dict().items()}
        prepared = prepare(sd, self.geometry, _round_up(self.geometry.input_channels, 4))
        self._prepared = _to_device(prepared, dev) if on_host else prepared
        self._prepared_sig = self._signature()
        self._plans.clear()
        self._weights_version = getattr(self, "_weights_version", 0) + 1  # captured CUDA graphs key on this (rollout.py)

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._prepared = None
        self._plans = {}
        self._weights_version = getattr(self, "_weights_version", 0) + 1
        return out

    # -- forward ---------------------------------------------------------------------------------------
    def _plan_for(self, x: torch.Tensor):
        """Validated input + the launch plan for its batch size / device (built on first use)."""
        if self.training:
            raise NotImplementedError("CrossFormerB200 implements the eval-mode forecast forward only: call .eval()")
        geo = self.geometry
        if x.dim() != 5 or tuple(x.shape[1:]) != geo.in_shape:
            raise ValueError(f"expected input [B, {', '.join(map(str, geo.in_shape))}], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("CrossFormerB200 has no CPU path: pass a CUDA tensor")
        if x.dtype != torch.float32:
            x = x.float()
        if self._prepared is None or self._prepared_sig != self._signature():
            self.refresh_weights()
        key = (int(x.shape[0]), x.device.index, self.exact_fp32)
        plan = self._plans.get(key)
        if plan is None:
            with torch.cuda.device(x.device):
                if self._domain is not None:
                    from .domain import DomainPlan

                    if x.shape[0] != 1:
                        raise ValueError("the domain-decomposed forward takes one state at a time (batch 1)")
                    dm = self._domain
                    plan = DomainPlan(geo, self._prepared, dm.domain_rank, dm.domain_world_size, x.device, dm.domain_group)
                else:
                    plan = _Plan(geo, self._prepared, int(x.shape[0]), x.device, not self.exact_fp32)
                self._plans[key] = plan
        return x.contiguous(), plan

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x, plan = self._plan_for(x)
        with torch.cuda.device(x.device):
            return plan.run(x)

    @torch.no_grad()
    def forward_fields(self, batch_input, pre, post=None) -> torch.Tensor:
        """The forecast step on the batch dict itself: ``batch_input`` is ``batch["input"]`` of the gen2
        pipeline (source -> var_key -> [B, n_levels, T, H, W] fp32 CUDA tensors, physical units), ``pre`` a
        ``pipeline.FusedPreblocks`` (normalisation + channel order, fused into the padding kernel: the concatenated input
        tensor is never built) and ``post`` an optional ``pipeline.FusedPostblocks`` (inverse scaling + tracer clamps in the
        epilogue of the output pass: the prediction comes back in physical units)."""
        if self._domain is not None:
            raise NotImplementedError("forward_fields runs on the single-GPU plan")
        tab = pre.tables(batch_input)
        probe = tab[8][0]
        if self.training:
            raise NotImplementedError("eval-mode forecast forward only: call .eval()")
        if self._prepared is None or self._prepared_sig != self._signature():
            self.refresh_weights()
        key = (tab[3], probe.device.index, self.exact_fp32)
        plan = self._plans.get(key)
        with torch.cuda.device(probe.device):
            if plan is None:
                plan = self._plans[key] = _Plan(self.geometry, self._prepared, tab[3], probe.device, not self.exact_fp32)
            return plan.run(None, fields=tab, post=post)


class WXFormerB200(CrossFormerB200):
    """Registry keys ``wxformer`` / ``wxformer_base`` of the reference (credit/models/wxformer/crossformer.py:623-904):
    same encoder, ZeroPad2d-wrapped cross-embed branches, PixelShuffle decoder.  Constructor keywords as there
    (``upsample_with_ps`` accepted and ignored, :669-673)."""

    VARIANT = "wxformer"

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.register_load_state_dict_pre_hook(WXFormerB200._legacy_keys_pre_hook)

    @staticmethod
    def _legacy_keys_pre_hook(module, state_dict, prefix, *args):
        """Checkpoints written before the ZeroPad2d wrap keep the cross-embed convs at ``convs.<i>.<suffix>``; the
        reference renames them to ``convs.<i>.1.<suffix>`` in a load_state_dict pre-hook and refuses checkpoints of the
        removed ConvTranspose2d decoder (wxformer/crossformer.py:239-310).  Same behaviour here."""
        import re

        if f"{prefix}up_block4.weight" in state_dict or f"{prefix}up_block4.weight_orig" in state_dict:
            raise RuntimeError("this checkpoint holds the ConvTranspose2d decoder (upsample_with_ps=False), which the "
                               "wxformer class no longer has: load it with type 'crossformer' / CrossFormerB200")
        pat = re.compile(r"^(layers\.\d+\.0\.convs\.\d+)\.(?!\d+\.)(.+)$")
        renamed = 0
        for key in [k for k in state_dict if k.startswith(prefix)]:
            m = pat.match(key[len(prefix):])
            if m is None:
                continue
            new = f"{prefix}{m.group(1)}.1.{m.group(2)}"
            if new not in state_dict:
                state_dict[new] = state_dict.pop(key)
                renamed += 1
        if renamed:
            logger.warning("Legacy CrossEmbedLayer checkpoint: remapped %d conv key(s) (convs.<i>.X -> convs.<i>.1.X)", renamed)


def register_with_credit(key: str = "crossformer_b200", message: Optional[str] = None):
    """Register under CREDIT's model registry (credit/models/__init__.py:128-161) so ``type: crossformer_b200`` selects
    ``CrossFormerB200`` from YAML; likewise ``wxformer_b200`` (PixelShuffle variant), ``fuxi_b200`` (FuXi) and
    ``crossformer-ensemble_b200`` (noise-injection variant).  Needs CREDIT importable."""
    from credit.models import register_model  # type: ignore

    from .ensemble import CrossFormerWithNoiseB200
    from .fuxi import FuxiB200

    register_model("wxformer_b200", "Loading the B200-native WXFormer (PixelShuffle decoder) forecast step ...")(WXFormerB200)
    register_model("fuxi_b200", "Loading the B200-native FuXi forecast step ...")(FuxiB200)
    register_model("crossformer-ensemble_b200", "Loading the B200-native CrossFormer with noise injection ...")(
        CrossFormerWithNoiseB200)
    return register_model(key, message or "Loading the B200-native CrossFormer forecast step ...")(CrossFormerB200)
