"""Exact latitude-band decomposition of ONE FuXi forecast step over the GPUs of a box (BASELINE config #4).

The reference's ``credit/domain_parallel`` converts FuXi's convolutions to halo-exchanging ones and leaves the Swin stage
"local within the shard" (convert.py:100-105), which changes the arithmetic of every shifted block (windows that straddle a
band boundary would be cut).  This decomposition keeps the single-GPU arithmetic:

* a rank owns whole WINDOW ROWS of the (zero-padded) token grid, hence a band of token / patch / pixel rows nested in it;
* convolutions: one halo row (3x3; the stride-2 DownBlock conv needs only the row above), GroupNorm statistics = the bands'
  fp64 sums added in rank order — the same primitives as the WXFormer decomposition (domain.py);
* un-shifted Swin blocks are local; a shifted block (cyclic shift 3) needs the first 3 token rows of the rank below: their
  q, k, v rows travel up before the attention, the attention output of those rows travels back down (two 3-row puts per
  shifted block, cyclic between the last and the first rank; the -100 row mask lives on the last rank only);
* the un-patchify / bilinear pass takes one halo patch row of the head's output from each neighbour.

All exchanges are NVLink peer-memory puts + arrival counters (peer.py, csrc/wxf_peer.cu); ``model(x)`` keeps the reference
contract (full state in, full prediction out on every rank), ``Rollout.step`` keeps the state sharded between steps.
"""

from __future__ import annotations

from typing import List

import torch

from . import lib as _lib
from . import ops
from .domain import DomainPlan, _Comm, _shift_taps, _split
from .fuxi import FuxiGeometry, FuxiWeights, StackWeights, _round_up


class FuxiDomainPlan:
    """Launch plan of one rank.  ``peer``: a ``peer.PeerComm`` (or the in-process stand-in of the CPU tests)."""

    def __init__(self, geo: FuxiGeometry, wts: FuxiWeights, rank: int, world: int, device, group=None, peer=None):
        g = self.geo = geo
        self.batch = 1
        self.rank, self.world = rank, world
        self.comm = _Comm(rank, world, group) if peer is None or not getattr(peer, "fake", False) else None
        ws, sft = g.ws[0], g.shift[0]
        if g.gh % ws or g.gh // ws < world:
            raise NotImplementedError(f"{g.gh // ws} window rows cannot be split over {world} ranks")
        left, right, top, bottom = g.pad2d
        wr = _split(g.gh // ws, world)
        self.p = [(ws * wr[r], ws * wr[r + 1]) for r in range(world)]                       # padded token rows
        self.t = [(min(max(a - top, 0), g.th), min(max(b - top, 0), g.th)) for a, b in self.p]  # token rows
        if any(b <= a for a, b in self.t):
            raise NotImplementedError("a rank would own window-padding rows only")
        self.sft = sft
        d, B = g.dim, 1
        f32 = dict(device=device, dtype=torch.float32)
        f16 = dict(device=device, dtype=torch.float16)
        n_of = lambda r: self.t[r][1] - self.t[r][0]       # noqa: E731  token rows of a rank
        rp_of = lambda r: self.p[r][1] - self.p[r][0]      # noqa: E731  padded token rows of a rank
        n, rp = n_of(rank), rp_of(rank)
        self.n, self.rp = n, rp
        tok = g.patch_height * g.patch_width * wts.cp
        self.cp, self.tok = wts.cp, tok

        def shapes(r):
            nn, rr = n_of(r), rp_of(r)
            sh = {}
            for pl in ("hi", "lo"):
                sh[f"ep.{pl}"] = ((1 + 2 * nn, g.lon, d), torch.float16)        # cube embedding + LN, top halo row
                sh[f"d0p.{pl}"] = ((nn + 2, g.tw, d), torch.float16)            # DownBlock conv output, halo both sides
                sh[f"dtmp.{pl}"] = ((nn + 2, g.tw, d), torch.float16)
                sh[f"up.{pl}"] = ((2 * nn + 2, g.lon, d), torch.float16)        # UpBlock ConvT output
                sh[f"utmp.{pl}"] = ((2 * nn + 2, g.lon, d), torch.float16)
                sh[f"att.{pl}"] = ((rr + sft, g.gw, d), torch.float16)          # attention output (+ 3 rows of the rank below)
            sh["qkv"] = ((rr + sft, g.gw, 3 * d), torch.float32)                  # q, k, v (+ 3 rows of the rank below)
            sh["ytok"] = ((2 * nn + 2, g.lon, tok), torch.float32)                # head output, one halo patch row each side
            for i in range(4):
                sh[f"gn_slots{i}"] = ((world, g.num_groups * 2), torch.float64)
            return sh

        numel = lambda shp: int(torch.Size(shp).numel())  # noqa: E731
        own = shapes(rank)
        every = self._every = [shapes(r) for r in range(world)]
        biggest = {k: max((every[r][k][0] for r in range(world)), key=numel) for k in own}
        if peer is None:
            from .peer import PeerComm, _ITEMSIZE

            total = sum((numel(biggest[k]) * _ITEMSIZE[own[k][1]] + 255) // 256 * 256 + 256 for k in own) + 64 * 256
            peer = PeerComm(rank, world, group, total, device)
        self.peer = peer
        self._buf, self._off = {}, {}
        for k, (shp, dt) in own.items():
            self._buf[k], self._off[k] = peer.buffer(biggest[k], shp, dt)
        bf = self._buf
        # local buffers
        self.ld0 = _round_up(g.in_chans * g.frames, 8)
        self.xp = (torch.empty((1, g.h_pad, g.w_pad, self.ld0), **f16), torch.empty((1, g.h_pad, g.w_pad, self.ld0), **f16))
        self.e0 = torch.empty((2 * n, g.lon, d), **f32)       # cube embedding, later the UpBlock shortcut
        self.a = torch.empty((2 * n, g.lon, d), **f32)        # conv outputs awaiting GroupNorm
        self.d0 = torch.empty((n, g.tw, d), **f32)
        self.sc = torch.empty((n, g.tw, d), **f32)
        self.catp = (torch.empty((n * g.tw, 2 * d), **f16), torch.empty((n * g.tw, 2 * d), **f16))
        M = self.M = rp * g.gw
        self.x = torch.empty((M, d), **f32)
        self.xpl = (torch.empty((M, d), **f16), torch.empty((M, d), **f16))
        self.tbuf = torch.empty((M, d), **f32)
        self.hid = (torch.empty((M, 4 * d), **f16), torch.empty((M, 4 * d), **f16))
        self.hp = (torch.empty((2 * n * g.lon, d), **f16), torch.empty((2 * n * g.lon, d), **f16))
        self.gn_sums = torch.empty((1, g.num_groups, 2), device=device, dtype=torch.float64)
        self.gn_stats = torch.empty((1, g.num_groups, 2), **f32)
        self.gn_scratch = torch.empty(ops.groupnorm_scratch_bytes(1, 2 * n * g.lon, d) // 4 + 4, **f32)
        # index lists: zero pad of the token band into the rank's window rows, and the crop back
        p0, p1 = self.p[rank]
        t0, t1 = self.t[rank]
        yy, xx = torch.meshgrid(torch.arange(p0, p1), torch.arange(g.gw), indexing="ij")
        inside = (yy >= top + t0) & (yy < top + t1) & (xx >= left) & (xx < left + g.tw)
        src = torch.where(inside, (yy - top - t0) * g.tw + (xx - left), torch.full_like(yy, -1))
        self.pad_idx = src.reshape(-1).to(device=device, dtype=torch.int32)
        ty, tx = torch.meshgrid(torch.arange(t0, t1), torch.arange(g.tw), indexing="ij")
        self.crop_idx = ((ty + top - p0) * g.gw + (tx + left)).reshape(-1).to(device=device, dtype=torch.int32)
        self._gn_calls = 0
        self.steps: List[tuple] = []
        self._keep: List[object] = []
        self._row_ranges()
        self._build(wts)

    # ---- plan construction helpers ---------------------------------------------------------------------------------
    def _add(self, fn, args, tag, flops=0.0, nbytes=0.0):
        self.steps.append((fn, args, tag, float(flops), float(nbytes)))

    def _conv_tc(self, in_hi, in_lo, wts, tag, **kw):
        desc = ops.make_conv_tc_desc(in_hi, in_lo, wts, **kw)
        m = kw["B"] * kw["Ho"] * kw["Wo"]
        self._add(ops.conv_f16x2_tc, (desc,), tag, 2.0 * m * wts.n * wts.t * wts.cin * wts.phases)

    def _gemm(self, a_hi, a_lo, wts, tag, **kw):
        desc = ops.make_gemm_desc(a_hi, a_lo, wts, **kw)
        self._add(ops.gemm_f16x2_tc, (desc,), tag, 2.0 * kw["M"] * wts.n * wts.k)

    def _rows_put(self, name, planes, src_row, n_rows, dst_rank, dst_row, site, slot):
        """Segments + signal of: rows [src_row, src_row + n_rows) of my buffer ``name`` -> rows [dst_row, ...) of the same
        buffer on ``dst_rank``; the arrival counter is ``site[slot]`` in the destination's arena."""
        P = self.peer
        segs = []
        for pl in (("hi", "lo") if planes else (None,)):
            key = name if pl is None else f"{name}.{pl}"
            t = self._buf[key]
            rb = t[0].numel() * t.element_size()
            segs.append((t.data_ptr() + src_row * rb, P.arena.base[dst_rank] + self._off[key] + dst_row * rb, n_rows * rb))
        return segs, P.sig(dst_rank, site, slot)

    def _halo(self, name, rows, up=True, down=True, planes=True):
        """Band buffer [1 + rows + 1, ...] (or [1 + rows] when only the row above is needed): my last interior row -> the top
        halo of the rank below (``down``), my first interior row -> the bottom halo of the rank above (``up``)."""
        P, me, world = self.peer, self.rank, self.world
        site = P.site()
        segs, sigs, waits = [], [], []
        if down and me < world - 1:
            s, sg = self._rows_put(name, planes, rows, 1, me + 1, 0, site, 0)
            segs += s
            sigs.append(sg)
        if up and me > 0:
            key = f"{name}.hi" if planes else name
            rows_prev = self._every[me - 1][key][0][0] - 2
            s, sg = self._rows_put(name, planes, 1, 1, me - 1, rows_prev + 1, site, 1)
            segs += s
            sigs.append(sg)
        if down and me > 0:
            waits.append(P.sig(me, site, 0))
        if up and me < world - 1:
            waits.append(P.sig(me, site, 1))
        self._add(self._put_wait, (segs, sigs, waits), "halo", 0, 0)

    def _put_wait(self, segs, sigs, waits):
        self.peer.put(segs, sigs)
        self.peer.wait(waits)

    def _gn_sums(self, x, hw_local, site):
        """GroupNorm, first half: the band's (sum, sum of squares) per group -> slot ``rank`` of every rank."""
        g, d, P = self.geo, self.geo.dim, self.peer
        sg, _slots, off = site
        G = g.num_groups
        ops.groupnorm_sums(x, d, self.gn_sums, self.gn_scratch, 1, hw_local, d, G)
        segs = [(self.gn_sums.data_ptr(), P.arena.base[r] + off + self.rank * 2 * G * 8, 2 * G * 8) for r in range(self.world)]
        P.put(segs, [P.sig(r, sg, self.rank) for r in range(self.world)])

    def _gn_apply(self, x, gamma, beta, res, out_f32, out_planes, hw_local, count, site):
        """GroupNorm, second half: statistics of the whole image (slots added in rank order), SiLU (+ residual)."""
        g, d, P = self.geo, self.geo.dim, self.peer
        sg, slots, _off = site
        G = g.num_groups
        P.wait_all(sg)
        P.sum_slots(slots, self.gn_sums, 2 * G)
        ops.groupnorm_stats_from_sums(self.gn_sums, self.gn_stats, 1, G, count)
        if out_f32 is not None:
            ops.groupnorm_apply(x, d, self.gn_stats, gamma, beta, res, d, out_f32, d, 1, hw_local, d, G)
        else:
            ops.groupnorm_apply_f16x2(x, d, self.gn_stats, gamma, beta, res, d, out_planes[0], out_planes[1], d, 0, 1, hw_local, d, G)

    def _gn_site(self):
        i = self._gn_calls
        self._gn_calls += 1
        return (self.peer.site(), self._buf[f"gn_slots{i}"], self._off[f"gn_slots{i}"])

    def _stack(self, sw: StackWeights, rows, width, x_f32, name_in, name_tmp, out_f32, out_planes, tag, hw_global):
        """2 x (conv3x3 + GroupNorm + SiLU) + skip on a band of ``rows`` rows; ``name_in`` / ``name_tmp``: halo'd plane buffers."""
        g, d = self.geo, self.geo.dim
        count = float(hw_global) * (d // g.num_groups)
        bi, bt = (self._buf[f"{name_in}.hi"], self._buf[f"{name_in}.lo"]), (self._buf[f"{name_tmp}.hi"], self._buf[f"{name_tmp}.lo"])
        a = self.a.view(-1)[: rows * width * d]
        self._halo(name_in, rows)
        self._conv_tc(bi[0], bi[1], _shift_taps(sw.convs[0], 1), f"{tag}_conv3x3", B=1, Hi=rows + 2, Wi=width, lda=d, Ho=rows,
                      Wo=width, out=a, ldc=d)
        site = self._gn_site()
        self._add(self._gn_sums, (a, rows * width, site), "groupnorm_silu", 0, 4.0 * rows * width * d)
        self._add(self._gn_apply, (a, sw.gn_w[0], sw.gn_b[0], None, None, (bt[0][1:], bt[1][1:]), rows * width, count, site),
                  "groupnorm_silu", 0, 8.0 * rows * width * d)
        self._halo(name_tmp, rows)
        self._conv_tc(bt[0], bt[1], _shift_taps(sw.convs[1], 1), f"{tag}_conv3x3", B=1, Hi=rows + 2, Wi=width, lda=d, Ho=rows,
                      Wo=width, out=a, ldc=d)
        site = self._gn_site()
        self._add(self._gn_sums, (a, rows * width, site), "groupnorm_silu", 0, 4.0 * rows * width * d)
        self._add(self._gn_apply, (a, sw.gn_w[1], sw.gn_b[1], x_f32, out_f32, out_planes, rows * width, count, site),
                  "groupnorm_silu", 0, 12.0 * rows * width * d)

    # ---- the plan ------------------------------------------------------------------------------------------------------
    def _build(self, wts: FuxiWeights):
        g, d, n, rp, M, me, world = self.geo, self.geo.dim, self.n, self.rp, self.M, self.rank, self.world
        add, bf, P, sft = self._add, self._buf, self.peer, self.sft
        t0 = self.t[me][0]
        # CubeEmbedding on the rank's pixel rows (no halo: kernel = stride), LayerNorm -> planes below the halo row
        px0 = 2 * t0 * g.patch_height
        self._conv_tc(self.xp[0][:, px0:], self.xp[1][:, px0:], wts.cube, "cube_embed", B=1, Hi=2 * n * g.patch_height,
                      Wi=g.w_pad, lda=self.ld0, Ho=2 * n, Wo=g.lon, out=self.e0, ldc=d)
        add(ops.layernorm_f16x2, (self.e0, d, bf["ep.hi"][1:], bf["ep.lo"][1:], d, wts.cube_g, wts.cube_b, 2 * n * g.lon, d),
            "layernorm", 0, 8.0 * 2 * n * g.lon * d)
        # DownBlock: conv3x3 stride 2 reads one patch row above the band
        self._halo("ep", 2 * n, up=False, down=True)
        self._conv_tc(bf["ep.hi"], bf["ep.lo"], _shift_taps(wts.down, 1), "down_conv", B=1, Hi=2 * n + 1, Wi=g.lon, lda=d, Ho=n,
                      Wo=g.tw, out=self.d0, ldc=d, out_hi=bf["d0p.hi"][1:], out_lo=bf["d0p.lo"][1:], ldh=d)
        self._stack(wts.down_stack, n, g.tw, self.d0, "d0p", "dtmp", self.sc, None, "down", g.th * g.tw)
        add(ops.split_f16x2, (self.sc, d, self.catp[0], self.catp[1], 2 * d, n * g.tw, d), "split", 0, 8.0 * n * g.tw * d)
        add(ops.gather_rows_ex, (self.sc, d, self.pad_idx, self.x, d, self.xpl[0], self.xpl[1], d, 0, M, d), "window_pad", 0,
            12.0 * M * d)
        # Swin-V2 stage on the rank's window rows
        L = g.ws[0] * g.ws[1]
        qkv, att = bf["qkv"], (bf["att.hi"], bf["att.lo"])
        prv, nxt = (me - 1) % world, (me + 1) % world
        for i, bw in enumerate(wts.blocks):
            shifted = any(g.block_shift(i))
            self._gemm(self.xpl[0], self.xpl[1], bw.qkv, "swin_qkv", M=M, lda=d, out=qkv, ldc=3 * d)
            if shifted:
                # my first `sft` rows of q, k, v complete the last window row of the rank above (cyclic)
                site = P.site()
                rows_prev = self.p[prv][1] - self.p[prv][0]
                segs, sg = self._rows_put("qkv", False, 0, sft, prv, rows_prev, site, 0)
                add(self._put_wait, (segs, [sg], [P.sig(me, site, 0)]), "shift_exchange", 0, 0)
                add(ops.swin_window_attention, (qkv[sft:], 3 * d, bw.bias, bw.logit_scale, att[0][sft:], att[1][sft:], None, d, 1,
                                                rp, g.gw, d, g.num_heads, g.ws, (0, g.shift[1]),
                                                g.shift[0] if me == world - 1 else 0), "swin_attention", 4.0 * M * L * d, 16.0 * M * d)
                # the attention output of those rows returns to the rank below
                site = P.site()
                segs, sg = self._rows_put("att", True, rp, sft, nxt, 0, site, 0)
                add(self._put_wait, (segs, [sg], [P.sig(me, site, 0)]), "shift_exchange", 0, 0)
            else:
                add(ops.swin_window_attention, (qkv, 3 * d, bw.bias, bw.logit_scale, att[0], att[1], None, d, 1, rp, g.gw, d,
                                                g.num_heads, g.ws, (0, 0), 0), "swin_attention", 4.0 * M * L * d, 16.0 * M * d)
            self._gemm(att[0], att[1], bw.proj, "swin_proj", M=M, lda=d, out=self.tbuf, ldc=d)
            add(ops.layernorm_residual, (self.tbuf, d, self.x, d, self.x, d, self.xpl[0], self.xpl[1], d, bw.n1_g, bw.n1_b, M, d),
                "layernorm_residual", 0, 16.0 * M * d)
            self._gemm(self.xpl[0], self.xpl[1], bw.fc1, "swin_fc1", M=M, lda=d, out_hi=self.hid[0], out_lo=self.hid[1], ldh=4 * d,
                       act=_lib.ACT_GELU)
            self._gemm(self.hid[0], self.hid[1], bw.fc2, "swin_fc2", M=M, lda=4 * d, out=self.tbuf, ldc=d)
            add(ops.layernorm_residual, (self.tbuf, d, self.x, d, self.x, d, self.xpl[0], self.xpl[1], d, bw.n2_g, bw.n2_b, M, d),
                "layernorm_residual", 0, 16.0 * M * d)
        add(ops.gather_rows_ex, (self.x, d, self.crop_idx, None, 0, self.catp[0], self.catp[1], 2 * d, d, n * g.tw, d), "window_crop",
            0, 8.0 * n * g.tw * d)
        # UpBlock: ConvTranspose k2 s2 (no halo), residual stack on the patch band
        self._conv_tc(self.catp[0], self.catp[1], wts.up, "up_convT", B=1, Hi=n, Wi=g.tw, lda=2 * d, Ho=n, Wo=g.tw, out=self.e0,
                      ldc=d, out_hi=bf["up.hi"][1:], out_lo=bf["up.lo"][1:], ldh=d)
        self._stack(wts.up_stack, 2 * n, g.lon, self.e0, "up", "utmp", None, self.hp, "up", g.lat * g.lon)
        # dense head, then one halo patch row of its output for the bilinear resize
        self._gemm(self.hp[0], self.hp[1], wts.head, "head", M=2 * n * g.lon, lda=d, out=bf["ytok"][1:], ldc=self.tok)
        self._halo("ytok", 2 * n, planes=False)

    # ---- boundary passes ------------------------------------------------------------------------------------------------
    def _row_ranges(self):
        """pad_rows / src_rows / out_rows of every rank (host bookkeeping, like DomainPlan._row_ranges)."""
        g, world = self.geo, self.world
        pt = g.padding.pad_lat[0] if g.padding.activate else 0
        H = g.image_height
        scale = float(torch.tensor(g.h_crop, dtype=torch.float32) / torch.tensor(g.h_out, dtype=torch.float32))
        o = torch.arange(g.h_out, dtype=torch.float64)
        src = (scale * (o + 0.5) - 0.5).float().clamp_min(0)
        y0 = src.floor().long().clamp_max(g.h_crop - 1)
        self.pad_rows, self.src_rows, self.out_rows = [], [], []
        ph2 = 2 * g.patch_height
        for r in range(world):
            a, b = ph2 * self.t[r][0], ph2 * self.t[r][1]
            self.pad_rows.append((a, b))
            rows = torch.arange(a, b)
            if g.padding.activate and g.padding.mode == "earth":
                sr = torch.where(rows < pt, pt - 1 - rows, torch.where(rows < pt + H, rows - pt, H - 1 - (rows - pt - H)))
            elif g.padding.activate:
                sr = (rows - pt).abs()
                sr = torch.where(sr >= H, 2 * (H - 1) - sr, sr)
            else:
                sr = rows
            self.src_rows.append((int(sr.min()), int(sr.max()) + 1))
            own = ((y0 + pt) >= a) & ((y0 + pt) < b)
            idx = own.nonzero().flatten()
            if idx.numel() == 0:
                self.out_rows.append((0, 0))
                continue
            lo, hi = int(idx[0]), int(idx[-1]) + 1
            if hi - lo != idx.numel():
                raise NotImplementedError("output rows are not contiguous per band")
            self.out_rows.append((lo, hi))
        if sum(b - a for a, b in self.out_rows) != g.h_out:
            raise NotImplementedError("output rows are not partitioned by the bands")

    def _pad(self, x):
        g = self.geo
        self.peer.advance()
        lat, lon, mode = ((g.padding.pad_lat, g.padding.pad_lon, g.padding.mode) if g.padding.activate
                          else ((0, 0), (0, 0), "earth"))
        a, b = self.pad_rows[self.rank]
        ops.pad_to_pixel_major_f16x2(x, lat, lon, mode, self.ld0, self.xp[0], self.xp[1], rows=(a, b - a))

    def _unpad(self, out):
        g = self.geo
        pt, pl = (g.padding.pad_lat[0], g.padding.pad_lon[0]) if g.padding.activate else (0, 0)
        lo, hi = self.out_rows[self.rank]
        ops.unpatchify_unpad_resize_to_nchw(self._buf["ytok"], out, 1, g.out_chans, self.cp, 2 * self.n + 2, g.lon, g.patch_height,
                                            g.patch_width, pt, pl, g.h_crop, g.w_crop, g.h_out, g.w_out, rows=(lo, hi - lo),
                                            lat0=2 * self.t[self.rank][0] - 1)

    exchange_rows = DomainPlan.exchange_rows

    def run_band(self, x: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        g = self.geo
        self._pad(x)
        for fn, args, _tag, _fl, _by in self.steps:
            fn(*args)
        if out is None:
            out = torch.empty((1, *g.out_shape), device=x.device, dtype=torch.float32)
        self._unpad(out)
        return out

    def run(self, x: torch.Tensor) -> torch.Tensor:
        """Full state in, full prediction out on every rank."""
        g = self.geo
        out = self.run_band(x)
        self.exchange_rows(out, g.out_chans, [(0, g.h_out)] * self.world)
        return out

    def run_profiled(self, x: torch.Tensor):
        recs = []

        def timed(tag, flops, nbytes, fn, *args):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(*args)
            e1.record()
            recs.append([tag, (e0, e1), flops, nbytes])

        timed("pad", 0.0, 0.0, self._pad, x)
        for fn, args, tag, fl, by in self.steps:
            timed(tag, fl, by, fn, *args)
        out = torch.empty((1, *self.geo.out_shape), device=x.device, dtype=torch.float32)
        timed("unpad_resize", 0.0, 0.0, self._unpad, out)
        torch.cuda.synchronize()
        return out, [(t, ev[0].elapsed_time(ev[1]), fl, by) for t, ev, fl, by in recs]
