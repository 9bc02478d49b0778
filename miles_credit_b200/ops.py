"""Thin tensor-level wrappers over the C-ABI entry points (one Python function per exported kernel).

PyTorch only supplies device memory and the current CUDA stream; every function enqueues exactly the
kernels of its entry point and raises ``RuntimeError`` on a non-zero status.  ``LAUNCHES`` counts the
kernels this package has enqueued (bench.py reports it as ``gpu_launches``).
"""

from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import lib as _lib
from .lib import WxfConvDesc, WxfConvTcDesc, WxfEnergyDesc, WxfGemmDesc, WxfToeplitzDesc, WxfWaterDesc
from .weights import ConvWeights

LAUNCHES = 0


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, name: str, dtype=torch.float32):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: miles_credit_b200 has no CPU path")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")


def pad_to_pixel_major(x: torch.Tensor, pad_lat, pad_lon, mode: str, ld: int, out: Optional[torch.Tensor] = None,
                       rows=None):
    """[B, C, T, H, W] -> padded pixel-major [B, Hp, Wp, ld] (channel c*T + t); ``rows`` = (first, count) of the padded
    rows to write (default: all)."""
    global LAUNCHES
    _req(x, "x")
    x = x.contiguous()
    b, c, t, h, w = x.shape
    hp, wp = h + pad_lat[0] + pad_lat[1], w + pad_lon[0] + pad_lon[1]
    if out is None:
        out = torch.empty((b, hp, wp, ld), device=x.device, dtype=torch.float32)
    m = _lib.PAD_EARTH if mode == "earth" else _lib.PAD_MIRROR
    r0, nr = rows if rows is not None else (0, hp)
    st = _lib.load().wxf_pad_to_pixel_major(x.data_ptr(), out.data_ptr(), b, c, t, h, w, pad_lat[0], pad_lat[1],
                                            pad_lon[0], pad_lon[1], m, ld, r0, nr, _stream())
    _lib.check(st, "wxf_pad_to_pixel_major")
    LAUNCHES += 1
    return out


def pad_to_pixel_major_f16x2(x: torch.Tensor, pad_lat, pad_lon, mode: str, ld: int, out_hi: torch.Tensor,
                             out_lo: torch.Tensor, rows=None):
    """[B, C, T, H, W] fp32 -> padded pixel-major fp16 hi/lo planes [B, Hp, Wp, ld]; ``rows`` = (first, count) of the
    padded rows to write (default: all)."""
    global LAUNCHES
    _req(x, "x")
    x = x.contiguous()
    b, c, t, h, w = x.shape
    m = _lib.PAD_EARTH if mode == "earth" else _lib.PAD_MIRROR
    r0, nr = rows if rows is not None else (0, h + pad_lat[0] + pad_lat[1])
    st = _lib.load().wxf_pad_to_pixel_major_f16x2(x.data_ptr(), out_hi.data_ptr(), out_lo.data_ptr(), b, c, t, h, w,
                                                  pad_lat[0], pad_lat[1], pad_lon[0], pad_lon[1], m, ld, r0, nr, _stream())
    _lib.check(st, "wxf_pad_to_pixel_major_f16x2")
    LAUNCHES += 1


def make_toeplitz_desc(in_hi: torch.Tensor, in_lo: torch.Tensor, wts, out: torch.Tensor, *, B: int, Hi: int, Wi: int,
                       lda: int, Ho: int, Wo: int, ldc: int, c_off: int = 0, oy_off: int = 0) -> WxfToeplitzDesc:
    """Descriptor of one stage-0 cross-embed branch; ``wts`` is a weights.ToeplitzWeights."""
    d = WxfToeplitzDesc()
    d.in_hi, d.in_lo = in_hi.data_ptr(), in_lo.data_ptr()
    d.w_hi, d.w_lo = wts.w_hi.data_ptr(), wts.w_lo.data_ptr()
    d.bias, d.out = _ptr(wts.bias), out.data_ptr()
    d.B, d.Hi, d.Wi, d.lda, d.Cin, d.cin_pad = B, Hi, Wi, lda, wts.cin, 64
    d.ch, d.kernel, d.pad = wts.ch, wts.kernel, wts.pad
    d.Ho, d.Wo, d.ldc, d.c_off = Ho, Wo, ldc, c_off
    d.w_scale_log2 = wts.scale_log2
    d.oy_off = oy_off
    return d


def cross_embed_toeplitz_tc(desc: WxfToeplitzDesc):
    global LAUNCHES
    st = _lib.load().wxf_cross_embed_toeplitz_tc(ctypes.byref(desc), _stream())
    _lib.check(st, "wxf_cross_embed_toeplitz_tc")
    LAUNCHES += 1


def layernorm(x: torch.Tensor, ldx: int, y: torch.Tensor, ldy: int, g: torch.Tensor, b: torch.Tensor, m: int, d: int,
              eps: float = 1e-5):
    global LAUNCHES
    st = _lib.load().wxf_layernorm(x.data_ptr(), ldx, y.data_ptr(), ldy, g.data_ptr(), b.data_ptr(), m, d, eps, _stream())
    _lib.check(st, "wxf_layernorm")
    LAUNCHES += 1


def layernorm_f16x2(x: torch.Tensor, ldx: int, y_hi: torch.Tensor, y_lo: torch.Tensor, ldh: int, g: torch.Tensor,
                    b: torch.Tensor, m: int, d: int, eps: float = 1e-5):
    """LayerNorm whose result is written as fp16 hi/lo operand planes."""
    global LAUNCHES
    st = _lib.load().wxf_layernorm_f16x2(x.data_ptr(), ldx, y_hi.data_ptr(), y_lo.data_ptr(), ldh, g.data_ptr(),
                                         b.data_ptr(), m, d, eps, _stream())
    _lib.check(st, "wxf_layernorm_f16x2")
    LAUNCHES += 1


def split_f16x2(x: torch.Tensor, ldx: int, hi: torch.Tensor, lo: torch.Tensor, ldh: int, m: int, d: int):
    global LAUNCHES
    st = _lib.load().wxf_split_f16x2(x.data_ptr(), ldx, hi.data_ptr(), lo.data_ptr(), ldh, m, d, _stream())
    _lib.check(st, "wxf_split_f16x2")
    LAUNCHES += 1


def make_gemm_desc(a_hi: torch.Tensor, a_lo: torch.Tensor, wts, *, M: int, lda: int, out: Optional[torch.Tensor] = None,
                   ldc: int = 0, c_off: int = 0, res: Optional[torch.Tensor] = None, ldr: int = 0, r_off: int = 0,
                   out_hi: Optional[torch.Tensor] = None, out_lo: Optional[torch.Tensor] = None, ldh: int = 0,
                   act: int = 0) -> WxfGemmDesc:
    """Descriptor of one tensor-core GEMM launch; ``wts`` is a weights.GemmWeights (fp16 hi/lo planes)."""
    d = WxfGemmDesc()
    d.a_hi, d.a_lo = a_hi.data_ptr(), a_lo.data_ptr()
    d.w_hi, d.w_lo = wts.w_hi.data_ptr(), wts.w_lo.data_ptr()
    d.bias, d.res, d.out = _ptr(wts.bias), _ptr(res), _ptr(out)
    d.out_hi, d.out_lo = _ptr(out_hi), _ptr(out_lo)
    d.M, d.N, d.K, d.lda = M, wts.n, wts.k, lda
    d.ldc, d.c_off, d.ldr, d.r_off, d.ldh = ldc, c_off, ldr, r_off, ldh
    d.act, d.w_scale_log2 = act, wts.scale_log2
    return d


def gemm_f16x2_tc(desc: WxfGemmDesc):
    global LAUNCHES
    st = _lib.load().wxf_gemm_f16x2_tc(ctypes.byref(desc), _stream())
    _lib.check(st, "wxf_gemm_f16x2_tc")
    LAUNCHES += 1


def make_conv_tc_desc(in_hi: torch.Tensor, in_lo: torch.Tensor, wts, *, B: int, Hi: int, Wi: int, lda: int, Ho: int,
                      Wo: int, out: Optional[torch.Tensor] = None, ldc: int = 0, c_off: int = 0,
                      res: Optional[torch.Tensor] = None, ldr: int = 0, r_off: int = 0,
                      out_hi: Optional[torch.Tensor] = None, out_lo: Optional[torch.Tensor] = None, ldh: int = 0,
                      h_off: int = 0, act: int = 0) -> WxfConvTcDesc:
    """Descriptor of one tensor-core implicit-GEMM convolution; ``wts`` is a weights.ConvTcWeights."""
    d = WxfConvTcDesc()
    d.in_hi, d.in_lo = in_hi.data_ptr(), in_lo.data_ptr()
    d.w_hi, d.w_lo = wts.w_hi.data_ptr(), wts.w_lo.data_ptr()
    d.taps = ctypes.addressof(wts.taps_host)
    d.bias, d.res, d.out = _ptr(wts.bias), _ptr(res), _ptr(out)
    d.out_hi, d.out_lo = _ptr(out_hi), _ptr(out_lo)
    d.B, d.Hi, d.Wi, d.lda, d.Cin, d.cin_pad = B, Hi, Wi, lda, wts.cin, wts.cin_pad
    d.N, d.T, d.stride = wts.n, wts.t, wts.stride
    d.Ho, d.Wo = Ho, Wo
    d.phases, d.out_scale = wts.phases, wts.out_scale
    d.ldc, d.c_off, d.ldr, d.r_off, d.ldh, d.h_off = ldc, c_off, ldr, r_off, ldh, h_off
    d.act, d.w_scale_log2 = act, wts.scale_log2
    d.bias_phase_stride = wts.bias_phase_stride
    d._keep = wts  # the host tap table must outlive the descriptor
    return d


def conv_f16x2_tc(desc: WxfConvTcDesc):
    global LAUNCHES
    st = _lib.load().wxf_conv_f16x2_tc(ctypes.byref(desc), _stream())
    _lib.check(st, "wxf_conv_f16x2_tc")
    LAUNCHES += 1


def groupnorm_silu_f16x2(x: torch.Tensor, ldx: int, stats: torch.Tensor, scratch: torch.Tensor, gamma: torch.Tensor,
                         beta: torch.Tensor, res: Optional[torch.Tensor], ldr: int, y_hi: torch.Tensor,
                         y_lo: torch.Tensor, ldh: int, h_off: int, B: int, HW: int, C: int, G: int, eps: float = 1e-5):
    """GroupNorm statistics + normalise/affine/SiLU (+ residual), result as fp16 hi/lo operand planes."""
    global LAUNCHES
    L = _lib.load()
    st = L.wxf_groupnorm_stats(x.data_ptr(), ldx, stats.data_ptr(), scratch.data_ptr(), B, HW, C, G, eps, _stream())
    _lib.check(st, "wxf_groupnorm_stats")
    st = L.wxf_groupnorm_silu_f16x2(x.data_ptr(), ldx, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(res), ldr,
                                    y_hi.data_ptr(), y_lo.data_ptr(), ldh, h_off, B, HW, C, G, _stream())
    _lib.check(st, "wxf_groupnorm_silu_f16x2")
    LAUNCHES += 3


def window_attention_f16x2(qkv: torch.Tensor, ldq: int, bias_t: torch.Tensor, out_hi: torch.Tensor, out_lo: torch.Tensor,
                           ldh: int, B: int, H: int, W: int, d: int, dh: int, wsz: int, kind: int, scale: float):
    global LAUNCHES
    st = _lib.load().wxf_window_attention_f16x2(qkv.data_ptr(), ldq, bias_t.data_ptr(), out_hi.data_ptr(),
                                                out_lo.data_ptr(), ldh, B, H, W, d, dh, wsz, kind, scale, _stream())
    _lib.check(st, "wxf_window_attention_f16x2")
    LAUNCHES += 1


def attention_bias_tile(bias_t: torch.Tensor, W: int, wsz: int, kind: int) -> torch.Tensor:
    """[128 x 128] fp32 bias tile of one Attention layer for wxf_window_attention_tc (built once per plan)."""
    global LAUNCHES
    tile = torch.empty(128 * 128, device=bias_t.device, dtype=torch.float32)
    st = _lib.load().wxf_attention_bias_tile(bias_t.data_ptr(), tile.data_ptr(), W, wsz, kind, _stream())
    _lib.check(st, "wxf_attention_bias_tile")
    LAUNCHES += 1
    return tile


def window_attention_tc(qkv_hi: torch.Tensor, qkv_lo: torch.Tensor, ldq: int, bias_t: torch.Tensor, out_hi: torch.Tensor,
                        out_lo: torch.Tensor, ldh: int, B: int, H: int, W: int, d: int, dh: int, wsz: int, kind: int,
                        scale: float):
    """Window attention on the tensor cores: fp16 hi/lo planes of qkv in, planes of the attention output out.

    ``bias_t`` here is the tile returned by :func:`attention_bias_tile`."""
    global LAUNCHES
    st = _lib.load().wxf_window_attention_tc(qkv_hi.data_ptr(), qkv_lo.data_ptr(), ldq, bias_t.data_ptr(),
                                             out_hi.data_ptr(), out_lo.data_ptr(), ldh, B, H, W, d, dh, wsz, kind, scale,
                                             _stream())
    _lib.check(st, "wxf_window_attention_tc")
    LAUNCHES += 1


def make_conv_desc(inp: torch.Tensor, wts: ConvWeights, out: torch.Tensor, *, B: int, Hi: int, Wi: int, lda: int,
                   Ho: int, Wo: int, ldc: int, c_off: int = 0, res: Optional[torch.Tensor] = None, ldr: int = 0,
                   r_off: int = 0, act: int = 0, in_off: int = 0) -> WxfConvDesc:
    """Descriptor of one implicit-GEMM launch.  ``in_off``: element offset of the first input channel."""
    d = WxfConvDesc()
    d.inp = inp.data_ptr() + 4 * in_off
    d.w = wts.w.data_ptr()
    d.taps = wts.taps.data_ptr()
    d.bias = _ptr(wts.bias)
    d.res = _ptr(res)
    d.out = out.data_ptr()
    d.B, d.Hi, d.Wi, d.lda, d.Cin = B, Hi, Wi, lda, wts.cin
    d.N, d.T, d.stride = wts.n, wts.t, wts.stride
    d.Ho, d.Wo = Ho, Wo
    d.phases, d.out_scale = wts.phases, wts.out_scale
    d.ldc, d.c_off, d.ldr, d.r_off, d.act = ldc, c_off, ldr, r_off, act
    d.bias_phase_stride = wts.bias_phase_stride
    return d


def conv_igemm_f32(desc: WxfConvDesc):
    global LAUNCHES
    st = _lib.load().wxf_conv_igemm_f32(ctypes.byref(desc), _stream())
    _lib.check(st, "wxf_conv_igemm_f32")
    LAUNCHES += 1


def window_attention_f32(qkv: torch.Tensor, ldq: int, bias_t: torch.Tensor, out: torch.Tensor, ldo: int, B: int, H: int,
                         W: int, d: int, dh: int, wsz: int, kind: int, scale: float):
    global LAUNCHES
    st = _lib.load().wxf_window_attention_f32(qkv.data_ptr(), ldq, bias_t.data_ptr(), out.data_ptr(), ldo, B, H, W, d, dh,
                                              wsz, kind, scale, _stream())
    _lib.check(st, "wxf_window_attention_f32")
    LAUNCHES += 1


def groupnorm_scratch_bytes(B: int, HW: int, C: int) -> int:
    return int(_lib.load().wxf_groupnorm_scratch_bytes(B, HW, C))


def groupnorm_silu(x: torch.Tensor, ldx: int, stats: torch.Tensor, scratch: torch.Tensor, gamma: torch.Tensor,
                   beta: torch.Tensor, res: Optional[torch.Tensor], ldr: int, y: torch.Tensor, ldy: int, B: int, HW: int,
                   C: int, G: int, eps: float = 1e-5):
    """GroupNorm statistics + normalise/affine/SiLU (+ residual)."""
    global LAUNCHES
    L = _lib.load()
    st = L.wxf_groupnorm_stats(x.data_ptr(), ldx, stats.data_ptr(), scratch.data_ptr(), B, HW, C, G, eps, _stream())
    _lib.check(st, "wxf_groupnorm_stats")
    st = L.wxf_groupnorm_silu(x.data_ptr(), ldx, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(res), ldr,
                              y.data_ptr(), ldy, B, HW, C, G, _stream())
    _lib.check(st, "wxf_groupnorm_silu")
    LAUNCHES += 3


def groupnorm_sums(x: torch.Tensor, ldx: int, sums: torch.Tensor, scratch: torch.Tensor, B: int, HW: int, C: int, G: int):
    """Local (sum, sum of squares) per (image, group) as fp64 [B, G, 2] (all-reduced by the caller)."""
    global LAUNCHES
    st = _lib.load().wxf_groupnorm_sums(x.data_ptr(), ldx, sums.data_ptr(), scratch.data_ptr(), B, HW, C, G, _stream())
    _lib.check(st, "wxf_groupnorm_sums")
    LAUNCHES += 2


def groupnorm_stats_from_sums(sums: torch.Tensor, stats: torch.Tensor, B: int, G: int, count: float, eps: float = 1e-5):
    global LAUNCHES
    st = _lib.load().wxf_groupnorm_stats_from_sums(sums.data_ptr(), stats.data_ptr(), B, G, float(count), eps, _stream())
    _lib.check(st, "wxf_groupnorm_stats_from_sums")
    LAUNCHES += 1


def groupnorm_apply(x, ldx, stats, gamma, beta, res, ldr, y, ldy, B, HW, C, G):
    """normalise + affine + SiLU (+ residual) with given (mean, rstd), fp32 out."""
    global LAUNCHES
    st = _lib.load().wxf_groupnorm_silu(x.data_ptr(), ldx, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(res), ldr,
                                        y.data_ptr(), ldy, B, HW, C, G, _stream())
    _lib.check(st, "wxf_groupnorm_silu")
    LAUNCHES += 1


def groupnorm_apply_f16x2(x, ldx, stats, gamma, beta, res, ldr, y_hi, y_lo, ldh, h_off, B, HW, C, G):
    """normalise + affine + SiLU (+ residual) with given (mean, rstd), fp16 hi/lo planes out."""
    global LAUNCHES
    st = _lib.load().wxf_groupnorm_silu_f16x2(x.data_ptr(), ldx, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                              _ptr(res), ldr, y_hi.data_ptr(), y_lo.data_ptr(), ldh, h_off, B, HW, C, G,
                                              _stream())
    _lib.check(st, "wxf_groupnorm_silu_f16x2")
    LAUNCHES += 1


def gather_rows(src: torch.Tensor, ld_src: int, idx: torch.Tensor, dst: torch.Tensor, ld_dst: int, n: int, d: int):
    """dst[i, :d] = src[idx[i], :d] (fp32 rows; idx int32)."""
    global LAUNCHES
    if n == 0:
        return
    st = _lib.load().wxf_gather_rows(src.data_ptr(), ld_src, idx.data_ptr(), dst.data_ptr(), ld_dst, n, d, _stream())
    _lib.check(st, "wxf_gather_rows")
    LAUNCHES += 1


def unpad_resize_to_nchw(y: torch.Tensor, ld: int, out: torch.Tensor, B: int, C: int, Hd: int, Wd: int, top: int,
                         left: int, Hc: int, Wc: int, Ho: int, Wo: int, rows=None):
    """``rows`` = (first, count) of the output rows to write (default: all)."""
    global LAUNCHES
    o0, n_out = rows if rows is not None else (0, Ho)
    st = _lib.load().wxf_unpad_resize_to_nchw(y.data_ptr(), ld, out.data_ptr(), B, C, Hd, Wd, top, left, Hc, Wc, Ho, Wo,
                                              o0, n_out, _stream())
    _lib.check(st, "wxf_unpad_resize_to_nchw")
    LAUNCHES += 1


def copy_channels(dst: torch.Tensor, src: torch.Tensor, groups):
    """dst[:, d0:d0+n] = src[:, s0:s0+n] for (d0, s0, n) in groups; tensors are [B, C, ...] contiguous."""
    global LAUNCHES
    _req(dst, "dst")
    _req(src, "src")
    if not (dst.is_contiguous() and src.is_contiguous()):
        raise ValueError("copy_channels needs contiguous tensors")
    B, dc = dst.shape[:2]
    sc = src.shape[1]
    plane = dst[0, 0].numel()
    if src[0, 0].numel() != plane or src.shape[0] != B:
        raise ValueError("copy_channels: plane/batch mismatch")
    n = len(groups)
    arr = ctypes.c_int32 * n
    d0 = arr(*[g[0] for g in groups])
    s0 = arr(*[g[1] for g in groups])
    ln = arr(*[g[2] for g in groups])
    st = _lib.load().wxf_copy_channels(dst.data_ptr(), dc, src.data_ptr(), sc, B, plane, d0, s0, ln, n, _stream())
    _lib.check(st, "wxf_copy_channels")
    LAUNCHES += 1


# ---- FuXi (credit/models/fuxi.py): the entry points on top of the shared contraction / normalisation kernels ----------

def layernorm_residual(x: torch.Tensor, ldx: int, res: Optional[torch.Tensor], ldr: int, out: Optional[torch.Tensor], ldo: int,
                       out_hi: Optional[torch.Tensor], out_lo: Optional[torch.Tensor], ldh: int, g: torch.Tensor,
                       b: torch.Tensor, m: int, d: int, eps: float = 1e-5):
    """o = res + LayerNorm(x) * g + b, as fp32 and / or fp16 hi/lo planes (res-post-norm of a Swin-V2 block)."""
    global LAUNCHES
    st = _lib.load().wxf_layernorm_residual(x.data_ptr(), ldx, _ptr(res), ldr, _ptr(out), ldo, _ptr(out_hi), _ptr(out_lo), ldh,
                                            g.data_ptr(), b.data_ptr(), m, d, eps, _stream())
    _lib.check(st, "wxf_layernorm_residual")
    LAUNCHES += 1


def swin_window_attention(qkv: torch.Tensor, ldq: int, bias: torch.Tensor, logit_scale: torch.Tensor, out_hi, out_lo, out_f32,
                          ldh: int, B: int, H: int, W: int, d: int, heads: int, ws, shift, mask_shift_h: int = -1):
    """Swin-V2 scaled-cosine window attention of one block (cyclic shift, per-head bias and scale, -100 shift masks)."""
    global LAUNCHES
    st = _lib.load().wxf_swin_window_attention(qkv.data_ptr(), ldq, bias.data_ptr(), logit_scale.data_ptr(), _ptr(out_hi),
                                               _ptr(out_lo), _ptr(out_f32), ldh, B, H, W, d, heads, ws[0], ws[1], shift[0],
                                               shift[1], mask_shift_h, _stream())
    _lib.check(st, "wxf_swin_window_attention")
    LAUNCHES += 1


def gather_rows_ex(src: torch.Tensor, ld_src: int, idx: torch.Tensor, dst: Optional[torch.Tensor], ld_dst: int, hi, lo, ldh: int,
                   h_off: int, n: int, d: int):
    """dst[i, :d] = src[idx[i], :d] or 0 where idx[i] < 0; fp32 and / or fp16 hi/lo planes at column offset h_off."""
    global LAUNCHES
    if n == 0:
        return
    st = _lib.load().wxf_gather_rows_ex(src.data_ptr(), ld_src, idx.data_ptr(), _ptr(dst), ld_dst, _ptr(hi), _ptr(lo), ldh,
                                        h_off, n, d, _stream())
    _lib.check(st, "wxf_gather_rows_ex")
    LAUNCHES += 1


def unpatchify_unpad_resize_to_nchw(y: torch.Tensor, out: torch.Tensor, B: int, C: int, cp: int, Lat: int, Lon: int, ph: int,
                                    pw: int, top: int, left: int, Hc: int, Wc: int, Ho: int, Wo: int, rows=None, lat0: int = 0):
    """Token-major dense-head output -> un-patchify, crop, bilinear resize, NCHW (fuxi.py:484-498).  ``lat0``: first patch row
    the buffer holds (a latitude band with halo rows; 0 = the whole grid)."""
    global LAUNCHES
    o0, n_out = rows if rows is not None else (0, Ho)
    st = _lib.load().wxf_unpatchify_unpad_resize_to_nchw(y.data_ptr(), out.data_ptr(), B, C, cp, Lat, Lon, ph, pw, top, left, Hc,
                                                         Wc, Ho, Wo, o0, n_out, lat0, _stream())
    _lib.check(st, "wxf_unpatchify_unpad_resize_to_nchw")
    LAUNCHES += 1


def history_update(x: torch.Tensor, y: torch.Tensor, forcing: Optional[torch.Tensor], n_prog: int, n_dyn: int):
    """In-place rollout update of x [B, C, T, H, W] with a history window: slide the frames, newest frame from the prediction
    y [B, Cy, Ty, H, W] (prognostic channels) and ``forcing`` [B, n_dyn, 1, H, W] (None = carried)."""
    global LAUNCHES
    _req(x, "x")
    _req(y, "y")
    if not (x.is_contiguous() and y.is_contiguous() and (forcing is None or forcing.is_contiguous())):
        raise ValueError("history_update needs contiguous tensors")
    B, C, T = x.shape[:3]
    plane = x.shape[3] * x.shape[4]
    if y.shape[0] != B or y.shape[3] * y.shape[4] != plane:
        raise ValueError("history_update: prediction / state shape mismatch")
    st = _lib.load().wxf_history_update(x.data_ptr(), y.data_ptr(), _ptr(forcing), B, C, T, n_prog, n_dyn, y.shape[1],
                                        y.shape[2], plane, _stream())
    _lib.check(st, "wxf_history_update")
    LAUNCHES += 1


def preblock_pad_to_pixel_major(chan_table: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, B: int, C: int, T: int, H: int,
                                W: int, pad_lat, pad_lon, mode: str, ld: int, out: Optional[torch.Tensor] = None,
                                out_hi: Optional[torch.Tensor] = None, out_lo: Optional[torch.Tensor] = None, rows=None):
    """Normalise + concat + pad in one pass: ``chan_table`` is an int64 device tensor [B*C] of plane addresses ([T, H, W] fp32
    each), ``mean`` / ``std`` are [C] fp32.  Output like ``pad_to_pixel_major`` (``out``) or its f16x2 variant."""
    global LAUNCHES
    hp, wp = H + pad_lat[0] + pad_lat[1], W + pad_lon[0] + pad_lon[1]
    r0, nr = rows if rows is not None else (0, hp)
    m = _lib.PAD_EARTH if mode == "earth" else _lib.PAD_MIRROR
    st = _lib.load().wxf_preblock_pad_to_pixel_major(chan_table.data_ptr(), mean.data_ptr(), std.data_ptr(), _ptr(out),
                                                     _ptr(out_hi), _ptr(out_lo), B, C, T, H, W, pad_lat[0], pad_lat[1],
                                                     pad_lon[0], pad_lon[1], m, ld, r0, nr, _stream())
    _lib.check(st, "wxf_preblock_pad_to_pixel_major")
    LAUNCHES += 1


def unpad_resize_post_to_nchw(y: torch.Tensor, ld: int, out: torch.Tensor, B: int, C: int, Hd: int, Wd: int, top: int, left: int,
                              Hc: int, Wc: int, Ho: int, Wo: int, scale: torch.Tensor, shift: torch.Tensor, lo: torch.Tensor,
                              hi: torch.Tensor, rows=None):
    """``unpad_resize_to_nchw`` with the inverse scaling and the tracer clamps in its epilogue (per output channel)."""
    global LAUNCHES
    o0, n_out = rows if rows is not None else (0, Ho)
    st = _lib.load().wxf_unpad_resize_post_to_nchw(y.data_ptr(), ld, out.data_ptr(), B, C, Hd, Wd, top, left, Hc, Wc, Ho, Wo, o0,
                                                   n_out, scale.data_ptr(), shift.data_ptr(), lo.data_ptr(), hi.data_ptr(),
                                                   _stream())
    _lib.check(st, "wxf_unpad_resize_post_to_nchw")
    LAUNCHES += 1


_MASS_SCRATCH = {}


def dry_mass_sums(q: torch.Tensor, sp: torch.Tensor, area: torch.Tensor, da: torch.Tensor, db: torch.Tensor, rows=None):
    """q [B, L, H, W] and sp [B, H, W] (views of an NCHW state: contiguous planes) -> fp64 [B, 2] (GlobalMassFixer sums)."""
    global LAUNCHES
    _req(q, "q")
    _req(sp, "sp")
    B, L, H, W = q.shape
    if q.stride(3) != 1 or q.stride(2) != W or sp.stride(2) != 1 or sp.stride(1) != W:
        raise ValueError("dry_mass_sums needs contiguous [H, W] planes")
    r0, nr = rows if rows is not None else (0, H)
    key = (q.device, B)
    if key not in _MASS_SCRATCH:
        n = int(_lib.load().wxf_dry_mass_scratch_bytes(B))
        _MASS_SCRATCH[key] = torch.zeros(n, device=q.device, dtype=torch.uint8)
    sums = torch.empty((B, 2), device=q.device, dtype=torch.float64)
    st = _lib.load().wxf_dry_mass_sums(q.data_ptr(), q.stride(0), q.stride(1), sp.data_ptr(), sp.stride(0), area.data_ptr(),
                                       da.data_ptr(), db.data_ptr(), B, L, r0 * W, nr * W, sums.data_ptr(),
                                       _MASS_SCRATCH[key].data_ptr(), _stream())
    _lib.check(st, "wxf_dry_mass_sums")
    LAUNCHES += 1
    return sums


_BUDGET_SCRATCH = {}


def _budget_scratch(device, B: int) -> torch.Tensor:
    key = (device, B)
    if key not in _BUDGET_SCRATCH:
        _BUDGET_SCRATCH[key] = torch.zeros(int(_lib.load().wxf_budget_scratch_bytes(B)), device=device, dtype=torch.uint8)
    return _BUDGET_SCRATCH[key]


def _plane3(t: torch.Tensor, name: str):
    """[B, L, H, W] view with contiguous planes -> (pointer, batch stride, level stride)."""
    _req(t, name)
    if t.dim() != 4 or t.stride(3) != 1 or t.stride(2) != t.shape[3]:
        raise ValueError(f"{name}: needs a [B, L, H, W] view with contiguous [H, W] planes")
    return t.data_ptr(), t.stride(0), t.stride(1)


def _plane2(t: torch.Tensor, name: str):
    """[B, H, W] view with contiguous planes -> (pointer, batch stride)."""
    _req(t, name)
    if t.dim() != 3 or t.stride(2) != 1 or t.stride(1) != t.shape[2]:
        raise ValueError(f"{name}: needs a [B, H, W] view with contiguous [H, W] planes")
    return t.data_ptr(), t.stride(0)


def water_budget_sums(q_pred, sp_pred, q_in, sp_in, precip, evapor, area, coef_a, coef_b, n_seconds: float, rows=None):
    """GlobalWaterFixer sums (conservation.py:208-231): fp64 [B, 3] = (sum area dTWC/dt, sum area E flux, sum area P flux)."""
    global LAUNCHES
    B, L, H, W = q_pred.shape
    r0, nr = rows if rows is not None else (0, H)
    d = WxfWaterDesc()
    d.q_pred, d.q_pred_bs, d.q_pred_ls = _plane3(q_pred, "q_pred")
    d.q_in, d.q_in_bs, d.q_in_ls = _plane3(q_in, "q_in")
    d.sp_pred, d.sp_pred_bs = _plane2(sp_pred, "sp_pred")
    d.sp_in, d.sp_in_bs = _plane2(sp_in, "sp_in")
    d.precip, d.precip_bs = _plane2(precip, "precip")
    d.evapor, d.evapor_bs = _plane2(evapor, "evapor")
    d.area, d.coef_a, d.coef_b = area.data_ptr(), coef_a.data_ptr(), coef_b.data_ptr()
    d.p0, d.np, d.B, d.L, d.n_seconds = r0 * W, nr * W, B, L, float(n_seconds)
    sums = torch.empty((B, 3), device=q_pred.device, dtype=torch.float64)
    st = _lib.load().wxf_water_budget_sums(ctypes.byref(d), sums.data_ptr(), _budget_scratch(q_pred.device, B).data_ptr(), _stream())
    _lib.check(st, "wxf_water_budget_sums")
    LAUNCHES += 1
    return sums


def make_energy_desc(pred3, pred2, in3, sp_in, toa_down_in, gph_surf, area, coef_a, coef_b, n_seconds: float, rows=None):
    """pred3 = (T, q, U, V) prediction views [B, L, H, W] of ONE tensor (same strides), pred2 = (sp, toa_up_solar, toa_up_olr,
    surf_down_solar, surf_up_solar, surf_down_lw, surf_up_lw, surf_sh, surf_lh) [B, H, W] views of one tensor, in3 = (T, q, U,
    V) input views (same strides)."""
    B, L, H, W = pred3[0].shape
    r0, nr = rows if rows is not None else (0, H)
    d = WxfEnergyDesc()
    p3 = [_plane3(t, "pred3") for t in pred3]
    i3 = [_plane3(t, "in3") for t in in3]
    p2 = [_plane2(t, "pred2") for t in pred2]
    if len({(a[1], a[2]) for a in p3}) != 1 or len({(a[1], a[2]) for a in i3}) != 1 or len({a[1] for a in p2}) != 1:
        raise ValueError("energy fixer: the fields of a group must be channel views of one tensor (equal strides)")
    d.t_pred, d.q_pred, d.u_pred, d.v_pred = (a[0] for a in p3)
    d.pred3_bs, d.pred3_ls = p3[0][1], p3[0][2]
    (d.sp_pred, d.toa_up_solar, d.toa_up_olr, d.surf_down_solar, d.surf_up_solar, d.surf_down_lw, d.surf_up_lw, d.surf_sh,
     d.surf_lh) = (a[0] for a in p2)
    d.pred2_bs = p2[0][1]
    d.t_in, d.q_in, d.u_in, d.v_in = (a[0] for a in i3)
    d.in3_bs, d.in3_ls = i3[0][1], i3[0][2]
    d.sp_in, d.sp_in_bs = _plane2(sp_in, "sp_in")
    d.toa_down_in, d.toa_down_bs = _plane2(toa_down_in, "toa_down_in")
    d.gph_surf, d.area, d.coef_a, d.coef_b = gph_surf.data_ptr(), area.data_ptr(), coef_a.data_ptr(), coef_b.data_ptr()
    d.p0, d.np, d.B, d.L, d.n_seconds = r0 * W, nr * W, B, L, float(n_seconds)
    return d


def energy_budget_sums(desc: WxfEnergyDesc, device) -> torch.Tensor:
    """GlobalEnergyFixerUpDown sums (conservation.py:314-366): fp64 [B, 4] = (sum area R_T, sum area F_S, TE(t0), TE(t1))."""
    global LAUNCHES
    sums = torch.empty((desc.B, 4), device=device, dtype=torch.float64)
    st = _lib.load().wxf_energy_budget_sums(ctypes.byref(desc), sums.data_ptr(), _budget_scratch(device, desc.B).data_ptr(), _stream())
    _lib.check(st, "wxf_energy_budget_sums")
    LAUNCHES += 1
    return sums


def energy_fix_temperature(desc: WxfEnergyDesc, ratio: torch.Tensor):
    """T_pred <- (E_level(t1) * ratio - E_qgk(t1)) / CP(t1), in place (conservation.py:368-372)."""
    global LAUNCHES
    st = _lib.load().wxf_energy_fix_temperature(ctypes.byref(desc), ratio.data_ptr(), _stream())
    _lib.check(st, "wxf_energy_fix_temperature")
    LAUNCHES += 1


def scale_planes(x: torch.Tensor, ratio: torch.Tensor):
    """x[b] *= ratio[b] in place for a [B, H, W] view with contiguous planes."""
    global LAUNCHES
    _req(x, "x")
    B, H, W = x.shape
    if x.stride(2) != 1 or x.stride(1) != W:
        raise ValueError("scale_planes needs contiguous [H, W] planes")
    st = _lib.load().wxf_scale_planes(x.data_ptr(), x.stride(0), H * W, ratio.data_ptr(), B, _stream())
    _lib.check(st, "wxf_scale_planes")
    LAUNCHES += 1


# ---- ensemble noise injection (crossformer_ensemble.py) ------------------------------------------------------------------------

def noise_coef(latent: Optional[torch.Tensor], W: torch.Tensor, bias: torch.Tensor, mod: torch.Tensor, factor: torch.Tensor,
               coef: torch.Tensor, B: int, C: int, D: int, seed: int, step_counter: Optional[torch.Tensor], site: int):
    global LAUNCHES
    st = _lib.load().wxf_noise_coef(_ptr(latent), W.data_ptr(), bias.data_ptr(), mod.data_ptr(), factor.data_ptr(), coef.data_ptr(),
                                    B, C, D, seed, _ptr(step_counter), site, _stream())
    _lib.check(st, "wxf_noise_coef")
    LAUNCHES += 1


def noise_inject(x: torch.Tensor, ldx: int, out: Optional[torch.Tensor], ldo: int, out_hi, out_lo, ldh: int, h_off: int,
                 coef: torch.Tensor, eps: Optional[torch.Tensor], B: int, HW: int, C: int, seed: int,
                 step_counter: Optional[torch.Tensor], site: int):
    global LAUNCHES
    st = _lib.load().wxf_noise_inject(x.data_ptr(), ldx, _ptr(out), ldo, _ptr(out_hi), _ptr(out_lo), ldh, h_off, coef.data_ptr(),
                                      _ptr(eps), B, HW, C, seed, _ptr(step_counter), site, _stream())
    _lib.check(st, "wxf_noise_inject")
    LAUNCHES += 1


def noise_step_advance(step_counter: torch.Tensor):
    global LAUNCHES
    st = _lib.load().wxf_noise_step_advance(step_counter.data_ptr(), _stream())
    _lib.check(st, "wxf_noise_step_advance")
    LAUNCHES += 1
