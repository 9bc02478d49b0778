"""Test double of the C-ABI library: executes the *documented semantics* of every entry point of
include/wxformer_b200.h on CPU memory, straight from raw pointers and dims (numpy views over ctypes).

It exists so the host logic (launch plan, descriptors, weight re-layout, tap tables, buffer aliasing) can be
checked in the CPU test-suite without a GPU.  It lives under tests/ and is never importable from the product:
``miles_credit_b200`` has no CPU path.
"""
import ctypes
import math

import numpy as np
import torch


def _arr(ptr, n):
    if not ptr:
        return None
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=(int(n),))


def _iarr(ptr, n):
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_int32)), shape=(int(n),))


def _t(a):
    return torch.from_numpy(a)


class EmulatedLib:
    """Same call signatures as the ctypes handle returned by ``miles_credit_b200.lib.load()``."""

    def __init__(self):
        self.calls = []
        # precision study only (tools/precision_table.py): "alo" drops the A_lo W_hi pass (activations rounded to fp16),
        # "wlo" drops the A_hi W_lo pass (weights rounded to fp16) of the GEMM entry point
        self.drop = set()

    @staticmethod
    def _no_overlap(what, reads, writes):
        """A contraction reads a halo of its input while other CTAs write the output: on the GPU an input byte range that
        overlaps an output byte range is a cross-CTA race (the emulator accumulates out of place and would hide it)."""
        for rn, r0, r1 in reads:
            for wn, w0, w1 in writes:
                if r0 and w0 and r0 < w1 and w0 < r1:
                    raise AssertionError(f"{what}: input {rn} [{r0:#x}, {r1:#x}) overlaps output {wn} [{w0:#x}, {w1:#x})")

    def wxf_abi_version(self):
        return 17

    def wxf_last_error(self):
        return b"emulator"

    def wxf_pad_to_pixel_major(self, x, xp, B, C, T, H, W, pt, pb, pl, pr, mode, ld, row0, nrows, stream):
        self.calls.append("pad")
        xs = _t(_arr(x, B * C * T * H * W)).view(B, C * T, H, W)
        Hp, Wp = H + pt + pb, W + pl + pr
        assert 0 <= row0 and row0 + nrows <= Hp
        out = _t(_arr(xp, B * Hp * Wp * ld)).view(B, Hp, Wp, ld)
        out[:, row0: row0 + nrows].zero_()
        for r in range(row0, row0 + nrows):
            roll = False
            if mode == 0:
                if r < pt:
                    sr, roll = pt - 1 - r, True
                elif r < pt + H:
                    sr = r - pt
                else:
                    sr, roll = H - 1 - (r - pt - H), True
            else:
                sr = abs(r - pt)
                if sr >= H:
                    sr = 2 * (H - 1) - sr
            j = (torch.arange(Wp) - pl) % W
            if roll:
                j = (j - W // 2) % W
            out[:, r, :, : C * T] = xs[:, :, sr, :][:, :, j].permute(0, 2, 1)
        return 0

    def wxf_pad_to_pixel_major_f16x2(self, x, xp_hi, xp_lo, B, C, T, H, W, pt, pb, pl, pr, mode, ld, row0, nrows,
                                     stream):
        Hp, Wp = H + pt + pb, W + pl + pr
        tmp = np.zeros(B * Hp * Wp * ld, dtype=np.float32)
        self.wxf_pad_to_pixel_major(x, tmp.ctypes.data, B, C, T, H, W, pt, pb, pl, pr, mode, ld, row0, nrows, stream)
        hi, lo = self._split(torch.from_numpy(tmp))
        sel = slice(row0, row0 + nrows)  # only the requested rows are written
        self._harr(xp_hi, tmp.size).view(B, Hp, Wp, ld)[:, sel].copy_(hi.view(B, Hp, Wp, ld)[:, sel])
        self._harr(xp_lo, tmp.size).view(B, Hp, Wp, ld)[:, sel].copy_(lo.view(B, Hp, Wp, ld)[:, sel])
        return 0

    def wxf_cross_embed_toeplitz_tc(self, dref, stream):
        d = dref._obj
        self.calls.append("toeplitz")
        k, p_, ch, J = d.kernel, d.pad, d.ch, d.kernel // 2
        N = J * ch
        if k % 2 or N > 256 or ch % 4 or d.cin_pad != 64 or d.Cin > 64:
            return -3
        B, Hi, Wi, lda, Cin, Ho, Wo = d.B, d.Hi, d.Wi, d.lda, d.Cin, d.Ho, d.Wo
        n_in = ((B * Hi - 1) * Wi + Wi - 1) * lda + Cin
        shape, strides = (B, Hi, Wi, Cin), (Hi * Wi * lda, Wi * lda, lda, 1)
        x_hi = self._harr(d.in_hi, n_in).as_strided(shape, strides).double()
        x_lo = self._harr(d.in_lo, n_in).as_strided(shape, strides).double()
        w_hi = self._harr(d.w_hi, N * 2 * k * 64).view(N, 2 * k, 64)[..., :Cin].double()
        w_lo = self._harr(d.w_lo, N * 2 * k * 64).view(N, 2 * k, 64)[..., :Cin].double()
        Mx = Wo + J - 1
        P = torch.zeros(B, Ho, Mx, N, dtype=torch.float64)
        oy, mm = torch.arange(Ho) * 2, torch.arange(Mx) * 2
        for ky in range(k):
            for r in range(2):
                iy, ix = oy + ky - p_ + 2 * d.oy_off, mm + r - p_
                mask = ((iy >= 0) & (iy < Hi))[:, None] & ((ix >= 0) & (ix < Wi))[None, :]
                gh = x_hi[:, iy.clamp(0, Hi - 1)][:, :, ix.clamp(0, Wi - 1)] * mask[None, :, :, None]
                gl = x_lo[:, iy.clamp(0, Hi - 1)][:, :, ix.clamp(0, Wi - 1)] * mask[None, :, :, None]
                ks = ky * 2 + r
                P += gh @ w_lo[:, ks].t() + gl @ w_hi[:, ks].t() + gh @ w_hi[:, ks].t()
        P = P.float() * (2.0 ** -d.w_scale_log2)
        out_v = torch.zeros(B, Ho, Wo, ch)
        if d.bias:
            out_v += _t(_arr(d.bias, ch))
        for j in range(J):
            out_v += P[:, :, j: j + Wo, j * ch: (j + 1) * ch]
        n_out = ((B * Ho - 1) * Wo + Wo - 1) * d.ldc + d.c_off + ch
        _t(_arr(d.out, n_out)).as_strided((B, Ho, Wo, ch), (Ho * Wo * d.ldc, Wo * d.ldc, d.ldc, 1), d.c_off).copy_(out_v)
        return 0

    def wxf_layernorm(self, x, ldx, y, ldy, g, b, M, d, eps, stream):
        self.calls.append("layernorm")
        xs = _t(_arr(x, (M - 1) * ldx + d)).as_strided((M, d), (ldx, 1))
        ys = _t(_arr(y, (M - 1) * ldy + d)).as_strided((M, d), (ldy, 1))
        gg, bb = _t(_arr(g, d)), _t(_arr(b, d))
        mean = xs.mean(1, keepdim=True)
        var = ((xs - mean) ** 2).mean(1, keepdim=True)
        ys.copy_((xs - mean) / (var + eps).sqrt() * gg + bb)
        return 0

    @staticmethod
    def _harr(ptr, n):
        return torch.from_numpy(np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint16)), shape=(int(n),))
                                .view(np.float16))

    @staticmethod
    def _split(v):
        hi = v.clamp(-65504.0, 65504.0).half()
        lo = (v - hi.float()).half()
        return hi, lo

    def wxf_layernorm_f16x2(self, x, ldx, y_hi, y_lo, ldh, g, b, M, d, eps, stream):
        self.calls.append("layernorm_f16x2")
        xs = _t(_arr(x, (M - 1) * ldx + d)).as_strided((M, d), (ldx, 1))
        gg, bb = _t(_arr(g, d)), _t(_arr(b, d))
        mean = xs.mean(1, keepdim=True)
        var = ((xs - mean) ** 2).mean(1, keepdim=True)
        hi, lo = self._split((xs - mean) / (var + eps).sqrt() * gg + bb)
        self._harr(y_hi, (M - 1) * ldh + d).as_strided((M, d), (ldh, 1)).copy_(hi)
        self._harr(y_lo, (M - 1) * ldh + d).as_strided((M, d), (ldh, 1)).copy_(lo)
        return 0

    def wxf_split_f16x2(self, x, ldx, hi_p, lo_p, ldh, M, d, stream):
        self.calls.append("split")
        xs = _t(_arr(x, (M - 1) * ldx + d)).as_strided((M, d), (ldx, 1))
        hi, lo = self._split(xs)
        self._harr(hi_p, (M - 1) * ldh + d).as_strided((M, d), (ldh, 1)).copy_(hi)
        self._harr(lo_p, (M - 1) * ldh + d).as_strided((M, d), (ldh, 1)).copy_(lo)
        return 0

    def wxf_gemm_f16x2_tc(self, dref, stream):
        d = dref._obj
        self.calls.append("gemm_tc")
        M, N, K = d.M, d.N, d.K
        a_hi = self._harr(d.a_hi, (M - 1) * d.lda + K).as_strided((M, K), (d.lda, 1)).double()
        a_lo = self._harr(d.a_lo, (M - 1) * d.lda + K).as_strided((M, K), (d.lda, 1)).double()
        w_hi = self._harr(d.w_hi, N * K).view(N, K).double()
        w_lo = self._harr(d.w_lo, N * K).view(N, K).double()
        acc = a_hi @ w_hi.t()
        if "wlo" not in self.drop:
            acc = acc + a_hi @ w_lo.t()
        if "alo" not in self.drop:
            acc = acc + a_lo @ w_hi.t()
        acc = acc.float()
        v = acc * (2.0 ** -d.w_scale_log2)
        if d.bias:
            v = v + _t(_arr(d.bias, N))
        if d.act == 1:
            v = 0.5 * v * (1 + torch.erf(v * 0.7071067811865476))
        if d.res:
            v = v + _t(_arr(d.res, (M - 1) * d.ldr + d.r_off + N)).as_strided((M, N), (d.ldr, 1), d.r_off)
        if d.out:
            _t(_arr(d.out, (M - 1) * d.ldc + d.c_off + N)).as_strided((M, N), (d.ldc, 1), d.c_off).copy_(v)
        if d.out_hi:
            hi, lo = self._split(v)
            self._harr(d.out_hi, (M - 1) * d.ldh + N).as_strided((M, N), (d.ldh, 1)).copy_(hi)
            self._harr(d.out_lo, (M - 1) * d.ldh + N).as_strided((M, N), (d.ldh, 1)).copy_(lo)
        return 0

    def wxf_conv_f16x2_tc(self, dref, stream):
        d = dref._obj
        self.calls.append("conv_tc")
        if d.N % 4 or d.cin_pad % 64 or d.T > 64:
            return -3
        B, Hi, Wi, lda, Cin, cp, N, T, s = d.B, d.Hi, d.Wi, d.lda, d.Cin, d.cin_pad, d.N, d.T, d.stride
        Ho, Wo, P, osc = d.Ho, d.Wo, d.phases, d.out_scale
        n_in = ((B * Hi - 1) * Wi + Wi - 1) * lda + Cin
        shape, strides = (B, Hi, Wi, Cin), (Hi * Wi * lda, Wi * lda, lda, 1)
        x_hi = self._harr(d.in_hi, n_in).as_strided(shape, strides).double()
        x_lo = self._harr(d.in_lo, n_in).as_strided(shape, strides).double()
        w_hi = self._harr(d.w_hi, P * N * T * cp).view(P, N, T, cp)[..., :Cin].double()
        w_lo = self._harr(d.w_lo, P * N * T * cp).view(P, N, T, cp)[..., :Cin].double()
        taps = _t(_iarr(d.taps, P * T * 2)).view(P, T, 2)
        Hout, Wout = Ho * osc, Wo * osc
        n_o = lambda ld, off: ((B * Hout - 1) * Wout + Wout - 1) * ld + off + N  # noqa: E731
        self._no_overlap("conv_tc", [("in_hi", d.in_hi or 0, (d.in_hi or 0) + 2 * n_in),
                                     ("in_lo", d.in_lo or 0, (d.in_lo or 0) + 2 * n_in)],
                         [("out", (d.out or 0) and d.out + 4 * d.c_off, (d.out or 0) + 4 * n_o(d.ldc, d.c_off)),
                          ("out_hi", (d.out_hi or 0) and d.out_hi + 2 * d.h_off, (d.out_hi or 0) + 2 * n_o(d.ldh, d.h_off)),
                          ("out_lo", (d.out_lo or 0) and d.out_lo + 2 * d.h_off, (d.out_lo or 0) + 2 * n_o(d.ldh, d.h_off))])

        def view(ptr, ld, off, harr=False):
            n_el = ((B * Hout - 1) * Wout + Wout - 1) * ld + off + N
            base = self._harr(ptr, n_el) if harr else _t(_arr(ptr, n_el))
            return base.as_strided((B, Hout, Wout, N), (Hout * Wout * ld, Wout * ld, ld, 1), off)

        res = view(d.res, d.ldr, d.r_off).clone() if d.res else None
        bias_all = _t(_arr(d.bias, N + (P - 1) * d.bias_phase_stride)) if d.bias else None
        oy, ox = torch.arange(Ho) * s, torch.arange(Wo) * s
        for z in range(P):
            bias = bias_all[z * d.bias_phase_stride: z * d.bias_phase_stride + N] if bias_all is not None else None
            acc = torch.zeros(B, Ho, Wo, N, dtype=torch.float64)
            for t in range(T):
                dy, dx = int(taps[z, t, 0]), int(taps[z, t, 1])
                iy, ix = oy + dy, ox + dx
                mask = ((iy >= 0) & (iy < Hi))[:, None] & ((ix >= 0) & (ix < Wi))[None, :]
                gh = x_hi[:, iy.clamp(0, Hi - 1)][:, :, ix.clamp(0, Wi - 1)] * mask[None, :, :, None]
                gl = x_lo[:, iy.clamp(0, Hi - 1)][:, :, ix.clamp(0, Wi - 1)] * mask[None, :, :, None]
                acc += gh @ w_lo[z, :, t].t() + gl @ w_hi[z, :, t].t() + gh @ w_hi[z, :, t].t()
            v = acc.float() * (2.0 ** -d.w_scale_log2)
            if bias is not None:
                v = v + bias
            if d.act == 1:
                v = 0.5 * v * (1 + torch.erf(v * 0.7071067811865476))
            py, px = z >> 1, z & 1
            sl = (slice(None), slice(py, None, osc), slice(px, None, osc)) if osc > 1 else (slice(None),) * 3
            if res is not None:
                v = v + res[sl]
            if d.out:
                view(d.out, d.ldc, d.c_off)[sl] = v
            if d.out_hi:
                hi, lo = self._split(v)
                view(d.out_hi, d.ldh, d.h_off, True)[sl] = hi
                view(d.out_lo, d.ldh, d.h_off, True)[sl] = lo
        return 0

    def wxf_groupnorm_silu_f16x2(self, x, ldx, stats, gamma, beta, res, ldr, y_hi, y_lo, ldh, h_off, B, HW, C, G, stream):
        tmp = np.zeros(B * HW * C, dtype=np.float32)
        self.wxf_groupnorm_silu(x, ldx, stats, gamma, beta, res, ldr, tmp.ctypes.data, C, B, HW, C, G, stream)
        hi, lo = self._split(torch.from_numpy(tmp).view(B * HW, C))
        n_el = (B * HW - 1) * ldh + h_off + C
        self._harr(y_hi, n_el).as_strided((B * HW, C), (ldh, 1), h_off).copy_(hi)
        self._harr(y_lo, n_el).as_strided((B * HW, C), (ldh, 1), h_off).copy_(lo)
        return 0

    def wxf_window_attention_f16x2(self, qkv, ldq, biasT, out_hi, out_lo, ldh, B, H, W, d, dh, wsz, kind, scale, stream):
        M = B * H * W
        tmp = np.zeros(M * d, dtype=np.float32)
        rc = self.wxf_window_attention_f32(qkv, ldq, biasT, tmp.ctypes.data, d, B, H, W, d, dh, wsz, kind, scale, stream)
        if rc:
            return rc
        hi, lo = self._split(torch.from_numpy(tmp).view(M, d))
        self._harr(out_hi, (M - 1) * ldh + d).as_strided((M, d), (ldh, 1)).copy_(hi)
        self._harr(out_lo, (M - 1) * ldh + d).as_strided((M, d), (ldh, 1)).copy_(lo)
        return 0

    def wxf_attention_bias_tile(self, biasT, tile, W, wsz, kind, stream):
        """The emulator keeps the tile opaque: it stores the plain transposed bias in the first L*L entries."""
        L = wsz * wsz
        _t(_arr(tile, 128 * 128))[: L * L] = _t(_arr(biasT, L * L))
        return 0

    def wxf_window_attention_tc(self, qkv_hi, qkv_lo, ldq, biasT, out_hi, out_lo, ldh, B, H, W, d, dh, wsz, kind, scale,
                                stream):
        """f16x2 semantics: operands are hi+lo (22 bits), the lo*lo products are dropped, softmax in fp32."""
        self.calls.append("attention_tc")
        if dh != 32 or wsz * wsz > 128 or ldq % 8:
            return -3
        M = B * H * W
        n_el = (M - 1) * ldq + 3 * d
        hi = self._harr(qkv_hi, n_el).as_strided((B, H, W, 3 * d), (H * W * ldq, W * ldq, ldq, 1)).float()
        lo = self._harr(qkv_lo, n_el).as_strided((B, H, W, 3 * d), (H * W * ldq, W * ldq, ldq, 1)).float()
        L = wsz * wsz
        bias = _t(_arr(biasT, L * L)).view(L, L).t()
        nh, nw, heads = H // wsz, W // wsz, d // dh
        res = torch.zeros(B, H, W, d)
        for b in range(B):
            for gh in range(nh):
                for gw in range(nw):
                    if kind == 0:
                        ys, xs = gh * wsz + torch.arange(wsz), gw * wsz + torch.arange(wsz)
                    else:
                        ys, xs = torch.arange(wsz) * nh + gh, torch.arange(wsz) * nw + gw
                    th = hi[b][ys][:, xs].reshape(L, 3, heads, dh).double()
                    tl = lo[b][ys][:, xs].reshape(L, 3, heads, dh).double()
                    qh, kh, vh = (th[:, i].transpose(0, 1) for i in range(3))
                    ql, kl, vl = (tl[:, i].transpose(0, 1) for i in range(3))
                    s_ = (qh @ kl.transpose(1, 2) + ql @ kh.transpose(1, 2) + qh @ kh.transpose(1, 2)).float()
                    s_ = s_ * scale + bias
                    pe = torch.exp(s_ - s_.max(-1, keepdim=True).values)
                    ph = pe.half().float()
                    pl = (pe - ph).half().float()
                    o = (ph.double() @ vl + pl.double() @ vh + ph.double() @ vh).float() / pe.sum(-1, keepdim=True)
                    res[b, ys[:, None], xs[None, :]] = o.transpose(0, 1).reshape(wsz, wsz, d)
        oh, ol = self._split(res.view(M, d))
        self._harr(out_hi, (M - 1) * ldh + d).as_strided((M, d), (ldh, 1)).copy_(oh)
        self._harr(out_lo, (M - 1) * ldh + d).as_strided((M, d), (ldh, 1)).copy_(ol)
        return 0

    def wxf_conv_igemm_f32(self, dref, stream):
        d = dref._obj
        self.calls.append("conv")
        B, Hi, Wi, lda, Cin, N, T, s = d.B, d.Hi, d.Wi, d.lda, d.Cin, d.N, d.T, d.stride
        Ho, Wo, P, osc = d.Ho, d.Wo, d.phases, d.out_scale
        n_in = ((B * Hi - 1) * Wi + Wi - 1) * lda + Cin
        xin = _t(_arr(d.inp, n_in)).as_strided((B, Hi, Wi, Cin), (Hi * Wi * lda, Wi * lda, lda, 1))
        w = _t(_arr(d.w, P * N * T * Cin)).view(P, N, T, Cin)
        taps = _t(_iarr(d.taps, P * T * 2)).view(P, T, 2)
        Hout, Wout = Ho * osc, Wo * osc
        n_out = ((B * Hout - 1) * Wout + Wout - 1) * d.ldc + d.c_off + N
        if T > 1:  # a 1x1 convolution reads only the pixel it writes
            self._no_overlap("conv", [("in", d.inp, d.inp + 4 * n_in)], [("out", d.out + 4 * d.c_off, d.out + 4 * n_out)])
        out = _t(_arr(d.out, n_out)).as_strided((B, Hout, Wout, N), (Hout * Wout * d.ldc, Wout * d.ldc, d.ldc, 1),
                                                d.c_off)
        res = None
        if d.res:
            n_res = ((B * Hout - 1) * Wout + Wout - 1) * d.ldr + d.r_off + N
            res = _t(_arr(d.res, n_res)).as_strided((B, Hout, Wout, N), (Hout * Wout * d.ldr, Wout * d.ldr, d.ldr, 1),
                                                    d.r_off).clone()
        bias_all = _t(_arr(d.bias, N + (P - 1) * d.bias_phase_stride)) if d.bias else None
        oy = torch.arange(Ho) * s
        ox = torch.arange(Wo) * s
        for z in range(P):
            bias = bias_all[z * d.bias_phase_stride: z * d.bias_phase_stride + N] if bias_all is not None else None
            acc = torch.zeros(B, Ho, Wo, N, dtype=torch.float64)
            for t in range(T):
                dy, dx = int(taps[z, t, 0]), int(taps[z, t, 1])
                iy, ix = oy + dy, ox + dx
                vy = (iy >= 0) & (iy < Hi)
                vx = (ix >= 0) & (ix < Wi)
                g = xin[:, iy.clamp(0, Hi - 1)][:, :, ix.clamp(0, Wi - 1)].double()
                g = g * (vy[:, None] & vx[None, :])[None, :, :, None]
                acc += g @ w[z, :, t, :].double().t()
            v = acc.float()
            if bias is not None:
                v = v + bias
            if d.act == 1:
                v = 0.5 * v * (1 + torch.erf(v * 0.7071067811865476))
            py, px = z >> 1, z & 1
            if res is not None:
                v = v + res[:, py::osc, px::osc] if osc > 1 else v + res
            if osc > 1:
                out[:, py::osc, px::osc] = v
            else:
                out.copy_(v)
        return 0

    def wxf_window_attention_f32(self, qkv, ldq, biasT, outp, ldo, B, H, W, d, dh, wsz, kind, scale, stream):
        self.calls.append("attention")
        if dh != 32 or wsz * wsz > 128:
            return -3
        M = B * H * W
        q3 = _t(_arr(qkv, (M - 1) * ldq + 3 * d)).as_strided((B, H, W, 3 * d), (H * W * ldq, W * ldq, ldq, 1)).clone()
        out = _t(_arr(outp, (M - 1) * ldo + d)).as_strided((B, H, W, d), (H * W * ldo, W * ldo, ldo, 1))
        L = wsz * wsz
        bias = _t(_arr(biasT, L * L)).view(L, L).t()
        nh, nw, heads = H // wsz, W // wsz, d // dh
        for b in range(B):
            for gh in range(nh):
                for gw in range(nw):
                    if kind == 0:
                        ys = gh * wsz + torch.arange(wsz)
                        xs = gw * wsz + torch.arange(wsz)
                    else:
                        ys = torch.arange(wsz) * nh + gh
                        xs = torch.arange(wsz) * nw + gw
                    tok = q3[b][ys][:, xs].reshape(L, 3, heads, dh)
                    q, k, v = tok[:, 0].transpose(0, 1) * scale, tok[:, 1].transpose(0, 1), tok[:, 2].transpose(0, 1)
                    p = (q @ k.transpose(1, 2) + bias).softmax(-1)
                    o = (p @ v).transpose(0, 1).reshape(wsz, wsz, d)
                    out[b, ys[:, None], xs[None, :]] = o
        return 0

    def wxf_groupnorm_scratch_bytes(self, B, HW, C):
        return B * ((HW + 511) // 512) * C * 8

    def wxf_groupnorm_stats(self, x, ldx, stats, scratch, B, HW, C, G, eps, stream):
        self.calls.append("gn_stats")
        xs = _t(_arr(x, (B * HW - 1) * ldx + C)).as_strided((B, HW, G, C // G), (HW * ldx, ldx, C // G, 1)).double()
        st = _t(_arr(stats, B * G * 2)).view(B, G, 2)
        mean = xs.mean(dim=(1, 3))
        var = (xs * xs).mean(dim=(1, 3)) - mean * mean
        st[..., 0] = mean.float()
        st[..., 1] = (1.0 / (var + eps).sqrt()).float()
        return 0

    def wxf_groupnorm_sums(self, x, ldx, sums, scratch, B, HW, C, G, stream):
        self.calls.append("gn_sums")
        xs = _t(_arr(x, (B * HW - 1) * ldx + C)).as_strided((B, HW, G, C // G), (HW * ldx, ldx, C // G, 1)).double()
        out = torch.from_numpy(np.ctypeslib.as_array(ctypes.cast(sums, ctypes.POINTER(ctypes.c_double)), shape=(B * G * 2,)))
        out = out.view(B, G, 2)
        out[..., 0] = xs.sum(dim=(1, 3))
        out[..., 1] = (xs * xs).sum(dim=(1, 3))
        return 0

    def wxf_groupnorm_stats_from_sums(self, sums, stats, B, G, count, eps, stream):
        sm = torch.from_numpy(np.ctypeslib.as_array(ctypes.cast(sums, ctypes.POINTER(ctypes.c_double)), shape=(B * G * 2,)))
        sm = sm.view(B, G, 2)
        st = _t(_arr(stats, B * G * 2)).view(B, G, 2)
        mean = sm[..., 0] / count
        var = (sm[..., 1] / count - mean * mean).clamp_min(0)
        st[..., 0] = mean.float()
        st[..., 1] = (1.0 / (var + eps).sqrt()).float()
        return 0

    def wxf_gather_rows(self, src, ld_src, idx, dst, ld_dst, n, d, stream):
        self.calls.append("gather_rows")
        ii = _t(_iarr(idx, n)).long()
        n_src = int(ii.max()) + 1
        s_ = _t(_arr(src, (n_src - 1) * ld_src + d)).as_strided((n_src, d), (ld_src, 1))
        _t(_arr(dst, (n - 1) * ld_dst + d)).as_strided((n, d), (ld_dst, 1)).copy_(s_[ii])
        return 0

    def wxf_groupnorm_silu(self, x, ldx, stats, gamma, beta, res, ldr, y, ldy, B, HW, C, G, stream):
        self.calls.append("gn_silu")
        xs = _t(_arr(x, (B * HW - 1) * ldx + C)).as_strided((B, HW, C), (HW * ldx, ldx, 1))
        st = _t(_arr(stats, B * G * 2)).view(B, G, 2)
        mean = st[..., 0].repeat_interleave(C // G, dim=1)[:, None, :]
        rstd = st[..., 1].repeat_interleave(C // G, dim=1)[:, None, :]
        v = (xs - mean) * rstd * _t(_arr(gamma, C)) + _t(_arr(beta, C))
        v = v / (1 + torch.exp(-v))
        if res:
            v = v + _t(_arr(res, (B * HW - 1) * ldr + C)).as_strided((B, HW, C), (HW * ldr, ldr, 1))
        _t(_arr(y, (B * HW - 1) * ldy + C)).as_strided((B, HW, C), (HW * ldy, ldy, 1)).copy_(v)
        return 0

    def wxf_unpad_resize_to_nchw(self, y, ld, outp, B, C, Hd, Wd, top, left, Hc, Wc, Ho, Wo, o0, n_out, stream):
        self.calls.append("unpad_resize")
        ys = _t(_arr(y, (B * Hd * Wd - 1) * ld + C)).as_strided((B, Hd, Wd, C), (Hd * Wd * ld, Wd * ld, ld, 1))
        crop = ys[:, top: top + Hc, left: left + Wc].permute(0, 3, 1, 2)
        out = _t(_arr(outp, B * C * Ho * Wo)).view(B, C, Ho, Wo)

        def axis(n_in, n_out):
            scale = float(np.float32(n_in) / np.float32(n_out))
            src = (scale * (torch.arange(n_out, dtype=torch.float64) + 0.5) - 0.5).float()  # fused multiply-add
            src = src.clamp_min(0)
            i0 = src.floor().long().clamp_max(n_in - 1)
            i1 = (i0 + 1).clamp_max(n_in - 1)
            l1 = (src - i0).clamp(0, 1)
            return i0, i1, 1 - l1, l1

        y0, y1, ly0, ly1 = axis(Hc, Ho)
        x0, x1, lx0, lx1 = axis(Wc, Wo)
        top_ = crop[:, :, y0][..., x0] * lx0 + crop[:, :, y0][..., x1] * lx1
        bot_ = crop[:, :, y1][..., x0] * lx0 + crop[:, :, y1][..., x1] * lx1
        assert 0 <= o0 and o0 + n_out <= Ho
        out[:, :, o0: o0 + n_out].copy_((top_ * ly0[:, None] + bot_ * ly1[:, None])[:, :, o0: o0 + n_out])
        return 0

    def wxf_copy_channels(self, dst, dst_C, src, src_C, B, plane, d0, s0, ln, n, stream):
        self.calls.append("copy_channels")
        dd = _t(_arr(dst, B * dst_C * plane)).view(B, dst_C, plane)
        ss = _t(_arr(src, B * src_C * plane)).view(B, src_C, plane)
        for g in range(n):
            dd[:, d0[g]: d0[g] + ln[g]] = ss[:, s0[g]: s0[g] + ln[g]]
        return 0

    # ---- global conservation fixers (documented semantics: fp32 per-pixel terms, fp64 area-weighted sums) --------------

    @staticmethod
    def _f3(ptr, bs, ls, B, L, p0, n):
        a = _t(_arr(ptr, (B - 1) * bs + (L - 1) * ls + p0 + n))
        return a.as_strided((B, L, n), (bs, ls, 1), p0)

    @staticmethod
    def _f2(ptr, bs, B, p0, n):
        a = _t(_arr(ptr, (B - 1) * bs + p0 + n))
        return a.as_strided((B, n), (bs, 1), p0)

    @staticmethod
    def _dp(ca, cb, sp):
        pres = ca.view(1, -1, 1) + cb.view(1, -1, 1) * sp.unsqueeze(1)      # [B, L + 1, n]
        return pres.diff(dim=1)

    def wxf_budget_scratch_bytes(self, B):
        return 64

    def wxf_scale_planes(self, x, bstride, n, ratio, B, stream):
        self.calls.append("scale_planes")
        xs = self._f2(x, bstride, B, 0, n)
        xs *= _t(_arr(ratio, B)).view(B, 1)
        return 0

    def wxf_water_budget_sums(self, dref, sums, scratch, stream):
        self.calls.append("water_budget_sums")
        d = dref._obj
        B, L, p0, n = d.B, d.L, d.p0, d.np
        ca, cb, area = _t(_arr(d.coef_a, L + 1)), _t(_arr(d.coef_b, L + 1)), _t(_arr(d.area, p0 + n))[p0:]
        nsec = torch.tensor(d.n_seconds, dtype=torch.float32)

        def twc(q, sp):
            return (q * self._dp(ca, cb, sp)).sum(1) / 9.80665

        sp1, sp0 = self._f2(d.sp_pred, d.sp_pred_bs, B, p0, n), self._f2(d.sp_in, d.sp_in_bs, B, p0, n)
        dtwc = (twc(self._f3(d.q_pred, d.q_pred_bs, d.q_pred_ls, B, L, p0, n), sp1)
                - twc(self._f3(d.q_in, d.q_in_bs, d.q_in_ls, B, L, p0, n), sp0)) / nsec
        ef = self._f2(d.evapor, d.evapor_bs, B, p0, n) * 1000.0 / nsec
        pf = self._f2(d.precip, d.precip_bs, B, p0, n) * 1000.0 / nsec
        out = np.ctypeslib.as_array(ctypes.cast(sums, ctypes.POINTER(ctypes.c_double)), shape=(B, 3))
        for k, t in enumerate((dtwc, ef, pf)):
            out[:, k] = (t * area).double().sum(1).numpy()
        return 0

    def _energy_fields(self, d):
        B, L, p0, n = d.B, d.L, d.p0, d.np
        f3p = lambda ptr: self._f3(ptr, d.pred3_bs, d.pred3_ls, B, L, p0, n)  # noqa: E731
        f3i = lambda ptr: self._f3(ptr, d.in3_bs, d.in3_ls, B, L, p0, n)  # noqa: E731
        gph = _t(_arr(d.gph_surf, p0 + n))[p0:].view(1, 1, n)

        def level_energy(T, q, U, V):
            cp = (1 - q) * 1004.64 + q * 1810.0
            e_qgk = 2.501e6 * q + gph + 0.5 * (U**2 + V**2)
            return cp, e_qgk, cp * T + e_qgk

        return f3p, f3i, level_energy

    def wxf_energy_budget_sums(self, dref, sums, scratch, stream):
        self.calls.append("energy_budget_sums")
        d = dref._obj
        B, L, p0, n = d.B, d.L, d.p0, d.np
        f3p, f3i, level_energy = self._energy_fields(d)
        f2p = lambda ptr: self._f2(ptr, d.pred2_bs, B, p0, n)  # noqa: E731
        ca, cb, area = _t(_arr(d.coef_a, L + 1)), _t(_arr(d.coef_b, L + 1)), _t(_arr(d.area, p0 + n))[p0:]
        nsec = torch.tensor(d.n_seconds, dtype=torch.float32)
        e1 = level_energy(f3p(d.t_pred), f3p(d.q_pred), f3p(d.u_pred), f3p(d.v_pred))[2]
        e0 = level_energy(f3i(d.t_in), f3i(d.q_in), f3i(d.u_in), f3i(d.v_in))[2]
        te1 = (e1 * self._dp(ca, cb, f2p(d.sp_pred))).sum(1) / 9.80665
        te0 = (e0 * self._dp(ca, cb, self._f2(d.sp_in, d.sp_in_bs, B, p0, n))).sum(1) / 9.80665
        r_t = (self._f2(d.toa_down_in, d.toa_down_bs, B, p0, n) * nsec - f2p(d.toa_up_solar) * nsec - f2p(d.toa_up_olr) * nsec) / nsec
        f_s = (f2p(d.surf_down_solar) - f2p(d.surf_up_solar) + f2p(d.surf_down_lw) - f2p(d.surf_up_lw) + f2p(d.surf_sh)
               + f2p(d.surf_lh)) / nsec
        out = np.ctypeslib.as_array(ctypes.cast(sums, ctypes.POINTER(ctypes.c_double)), shape=(B, 4))
        for k, t in enumerate((r_t, f_s, te0, te1)):
            out[:, k] = (t * area).double().sum(1).numpy()
        return 0

    def wxf_energy_fix_temperature(self, dref, ratio, stream):
        self.calls.append("energy_fix_temperature")
        d = dref._obj
        f3p, _f3i, level_energy = self._energy_fields(d)
        T = f3p(d.t_pred)
        cp, e_qgk, e = level_energy(T, f3p(d.q_pred), f3p(d.u_pred), f3p(d.v_pred))
        T.copy_((e * _t(_arr(ratio, d.B)).view(-1, 1, 1) - e_qgk) / cp)
        return 0

    # ---- FuXi entry points (documented semantics of include/wxformer_b200.h) ------------------------------------------

    def wxf_layernorm_residual(self, x, ldx, res, ldr, out, ldo, out_hi, out_lo, ldh, g, b, M, d, eps, stream):
        self.calls.append("layernorm_residual")
        xs = _t(_arr(x, (M - 1) * ldx + d)).as_strided((M, d), (ldx, 1))
        gg, bb = _t(_arr(g, d)), _t(_arr(b, d))
        mean = xs.mean(1, keepdim=True)
        var = ((xs - mean) ** 2).mean(1, keepdim=True)
        v = (xs - mean) / (var + eps).sqrt() * gg + bb
        if res:
            v = v + _t(_arr(res, (M - 1) * ldr + d)).as_strided((M, d), (ldr, 1))
        if out:
            _t(_arr(out, (M - 1) * ldo + d)).as_strided((M, d), (ldo, 1)).copy_(v)
        if out_hi:
            hi, lo = self._split(v)
            self._harr(out_hi, (M - 1) * ldh + d).as_strided((M, d), (ldh, 1)).copy_(hi)
            self._harr(out_lo, (M - 1) * ldh + d).as_strided((M, d), (ldh, 1)).copy_(lo)
        return 0

    def wxf_swin_window_attention(self, qkv, ldq, bias, logit_scale, out_hi, out_lo, out_f32, ldh, B, H, W, d, heads, ws_h,
                                  ws_w, shift_h, shift_w, mask_shift_h, stream):
        if mask_shift_h < 0:
            mask_shift_h = shift_h
        self.calls.append("swin_attention")
        dh, L = d // heads, ws_h * ws_w
        q3 = _t(_arr(qkv, (B * H * W - 1) * ldq + 3 * d)).as_strided((B, H, W, 3 * d), (H * W * ldq, W * ldq, ldq, 1))
        bs = _t(_arr(bias, heads * L * L)).view(heads, L, L)
        sc = _t(_arr(logit_scale, heads))
        rolled = torch.roll(q3, shifts=(-shift_h, -shift_w), dims=(1, 2))
        nwy, nwx = H // ws_h, W // ws_w
        win = rolled.view(B, nwy, ws_h, nwx, ws_w, 3, heads, dh).permute(5, 0, 1, 3, 6, 2, 4, 7).reshape(3, B, nwy, nwx, heads, L, dh)
        q, k, v = win[0].double(), win[1].double(), win[2].double()
        qn = q / q.norm(dim=-1, keepdim=True).clamp_min(1e-12)
        kn = k / k.norm(dim=-1, keepdim=True).clamp_min(1e-12)
        s = (qn @ kn.transpose(-1, -2)) * sc.view(1, 1, 1, heads, 1, 1).double() + bs.view(1, 1, 1, heads, L, L).double()

        def region(n, ws, sh):
            r = torch.zeros(n, dtype=torch.long)
            if sh > 0:
                r[n - ws: n - sh] = 1
                r[n - sh:] = 2
            return r

        ry, rx = region(H, ws_h, mask_shift_h), region(W, ws_w, shift_w)
        rid = (3 * ry[:, None] + rx[None, :]).view(nwy, ws_h, nwx, ws_w).permute(0, 2, 1, 3).reshape(nwy, nwx, L)
        mask = torch.where(rid[..., :, None] != rid[..., None, :], -100.0, 0.0).double()
        s = s + mask.view(1, nwy, nwx, 1, L, L)
        o = (torch.softmax(s, dim=-1) @ v).float()                                   # [B, nwy, nwx, heads, L, dh]
        o = o.view(B, nwy, nwx, heads, ws_h, ws_w, dh).permute(0, 1, 4, 2, 5, 3, 6).reshape(B, H, W, d)
        o = torch.roll(o, shifts=(shift_h, shift_w), dims=(1, 2)).reshape(B * H * W, d)
        M = B * H * W
        if out_f32:
            _t(_arr(out_f32, (M - 1) * ldh + d)).as_strided((M, d), (ldh, 1)).copy_(o)
        else:
            hi, lo = self._split(o)
            self._harr(out_hi, (M - 1) * ldh + d).as_strided((M, d), (ldh, 1)).copy_(hi)
            self._harr(out_lo, (M - 1) * ldh + d).as_strided((M, d), (ldh, 1)).copy_(lo)
        return 0

    def wxf_gather_rows_ex(self, src, ld_src, idx, dst, ld_dst, hi_p, lo_p, ldh, h_off, n, d, stream):
        self.calls.append("gather_rows_ex")
        ii = _t(_iarr(idx, n)).long()
        n_src = int(ii.max()) + 1
        s_ = _t(_arr(src, (n_src - 1) * ld_src + d)).as_strided((n_src, d), (ld_src, 1))
        v = torch.where((ii >= 0)[:, None], s_[ii.clamp_min(0)], torch.zeros(()))
        if dst:
            _t(_arr(dst, (n - 1) * ld_dst + d)).as_strided((n, d), (ld_dst, 1)).copy_(v)
        if hi_p:
            hi, lo = self._split(v)
            self._harr(hi_p, (n - 1) * ldh + h_off + d).as_strided((n, d), (ldh, 1), h_off).copy_(hi)
            self._harr(lo_p, (n - 1) * ldh + h_off + d).as_strided((n, d), (ldh, 1), h_off).copy_(lo)
        return 0

    def wxf_unpatchify_unpad_resize_to_nchw(self, y, outp, B, C, cp, Lat, Lon, ph, pw, top, left, Hc, Wc, Ho, Wo, o0, n_out,
                                            lat0, stream):
        self.calls.append("unpatchify_resize")
        ys = _t(_arr(y, B * Lat * Lon * ph * pw * cp)).view(B, Lat, Lon, ph, pw, cp)[..., :C]
        band = ys.permute(0, 1, 3, 2, 4, 5).reshape(B, Lat * ph, Lon * pw, C)
        # place the band (patch rows [lat0, lat0 + Lat)) into a whole-grid image; rows outside it must not be read
        n_all = max(top + Hc, (lat0 + Lat) * ph)
        img = torch.full((B, n_all, Lon * pw, C), float("nan"))
        lo = max(lat0 * ph, 0)
        img[:, lo: (lat0 + Lat) * ph] = band[:, lo - lat0 * ph:]
        Lat = n_all // ph if n_all % ph == 0 else (n_all + ph - 1) // ph
        if img.shape[1] != Lat * ph:
            img = torch.cat([img, torch.full((B, Lat * ph - img.shape[1], Lon * pw, C), float("nan"))], dim=1)
        img = img.contiguous()
        # the rest is wxf_unpad_resize_to_nchw on a pixel-major image with ld = C
        keep = self.calls
        self.calls = []
        rc = self.wxf_unpad_resize_to_nchw(img.data_ptr(), C, outp, B, C, Lat * ph, Lon * pw, top, left, Hc, Wc, Ho, Wo, o0,
                                           n_out, stream)
        self.calls = keep
        return rc

    def wxf_history_update(self, x, y, forcing, B, C, T, n_prog, n_dyn, Cy, Ty, plane, stream):
        self.calls.append("history_update")
        xs = _t(_arr(x, B * C * T * plane)).view(B, C, T, plane)
        ys = _t(_arr(y, B * Cy * Ty * plane)).view(B, Cy, Ty, plane)
        new = xs.clone()
        new[:, :, : T - 1] = xs[:, :, 1:]
        new[:, :n_prog, T - 1] = ys[:, :n_prog, 0]
        if forcing:
            new[:, n_prog: n_prog + n_dyn, T - 1] = _t(_arr(forcing, B * n_dyn * plane)).view(B, n_dyn, plane)
        xs.copy_(new)
        return 0

    # ---- pre / post-blocks fused into the boundary passes -----------------------------------------------------------------

    def wxf_preblock_pad_to_pixel_major(self, chan, mean, stdv, xp, xp_hi, xp_lo, B, C, T, H, W, pt, pb, pl, pr, mode, ld, row0,
                                        nrows, stream):
        table = np.ctypeslib.as_array(ctypes.cast(chan, ctypes.POINTER(ctypes.c_int64)), shape=(B * C,))
        m, s = _t(_arr(mean, C)), _t(_arr(stdv, C)).clamp(min=1e-12)
        x = torch.empty(B, C, T, H, W)
        for b in range(B):
            for c in range(C):
                x[b, c] = (_t(_arr(int(table[b * C + c]), T * H * W)).view(T, H, W) - m[c]) / s[c]
        x = x.contiguous()
        if xp:
            return self.wxf_pad_to_pixel_major(x.data_ptr(), xp, B, C, T, H, W, pt, pb, pl, pr, mode, ld, row0, nrows, stream)
        return self.wxf_pad_to_pixel_major_f16x2(x.data_ptr(), xp_hi, xp_lo, B, C, T, H, W, pt, pb, pl, pr, mode, ld, row0, nrows,
                                                 stream)

    def wxf_unpad_resize_post_to_nchw(self, y, ld, outp, B, C, Hd, Wd, top, left, Hc, Wc, Ho, Wo, o0, n_out, scale, shift, lo, hi,
                                      stream):
        rc = self.wxf_unpad_resize_to_nchw(y, ld, outp, B, C, Hd, Wd, top, left, Hc, Wc, Ho, Wo, o0, n_out, stream)
        out = _t(_arr(outp, B * C * Ho * Wo)).view(B, C, Ho, Wo)[:, :, o0: o0 + n_out]
        v = out * _t(_arr(scale, C)).view(1, C, 1, 1) + _t(_arr(shift, C)).view(1, C, 1, 1)
        out.copy_(torch.minimum(torch.maximum(v, _t(_arr(lo, C)).view(1, C, 1, 1)), _t(_arr(hi, C)).view(1, C, 1, 1)))
        return rc

    # ---- ensemble noise injection (recorded draws only: the emulator has no generator) ---------------------------------------

    def wxf_noise_coef(self, latent, W, bias, mod, factor, coef, B, C, D, seed, step, site, stream):
        self.calls.append("noise_coef")
        assert latent, "the emulator needs the latent draw"
        lat = _t(_arr(latent, B * D)).view(B, D)
        style = lat @ _t(_arr(W, C * D)).view(C, D).t() + _t(_arr(bias, C))
        _t(_arr(coef, B * C)).view(B, C).copy_(_t(_arr(factor, 1)) * style * _t(_arr(mod, C)))
        return 0

    def wxf_noise_inject(self, x, ldx, out, ldo, out_hi, out_lo, ldh, h_off, coef, eps, B, HW, C, seed, step, site, stream):
        self.calls.append("noise_inject")
        assert eps, "the emulator needs the eps draw"
        M = B * HW
        xs = _t(_arr(x, (M - 1) * ldx + C)).as_strided((M, C), (ldx, 1))
        k = _t(_arr(coef, B * C)).view(B, 1, C).expand(B, HW, C).reshape(M, C)
        v = xs + _t(_arr(eps, M * C)).view(M, C) * k
        if out:
            _t(_arr(out, (M - 1) * ldo + C)).as_strided((M, C), (ldo, 1)).copy_(v)
        if out_hi:
            hi, lo = self._split(v)
            self._harr(out_hi, (M - 1) * ldh + h_off + C).as_strided((M, C), (ldh, 1), h_off).copy_(hi)
            self._harr(out_lo, (M - 1) * ldh + h_off + C).as_strided((M, C), (ldh, 1), h_off).copy_(lo)
        return 0

    def wxf_noise_step_advance(self, step, stream):
        return 0
