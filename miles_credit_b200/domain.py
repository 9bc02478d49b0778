"""Exact lat-lon domain decomposition of ONE forecast step over the GPUs of a box (torch.distributed, NCCL).

The reference's ``credit/domain_parallel`` shards equal latitude bands and leaves attention "local within the shard"
(convert.py:100-105), which changes the arithmetic of the dilated long attention (measured 1.2e-1 off its own
single-device forward).  This decomposition is exact (same arithmetic as the single-GPU plan up to fp32
reassociation) and uses two layouts:

* **band layout** (convolutions, GroupNorm, decoder): rank r owns a contiguous band of grid rows, nested across stages
  (band at stage s = 2x band at stage s+1), with one-row halos exchanged between neighbours
  (the reference's halo widths: tests/test_domain_parallel.py:66-99; stage 0 reads the replicated padded input).
* **unit layout** (the transformer stack of a stage): the windows (wy, wx) with fixed (wy mod A, wx mod B),
  A = H/(lws*gws), B = W/(lws*gws), form a *unit*: a (lws*gws) x (lws*gws) mini-image that is closed under BOTH the
  short windows and the dilated long groups, and inside which they are again the ordinary short / long patterns.  A rank
  owns whole units and runs the unchanged single-GPU kernels on them as a batch: no communication inside a stage.

Per step: 2 all-to-alls per stage (band <-> unit, the fp32 residual stream), 1-row halo exchanges in the decoder,
one all-reduce of (sum, sum-of-squares) per GroupNorm.  Boundary padding and un-pad + resize are sharded too: a rank pads
only the rows its stage-0 band reads and writes only its own rows of the prediction.  ``model(x)`` all-gathers those rows
(full prediction on every rank, the reference contract); the rollout (rollout.py) keeps the state sharded between steps
and exchanges only the halo rows the next step's padding needs.
"""

from __future__ import annotations

import logging
import os
from typing import List, Optional

import torch
import torch.distributed as dist

from . import lib as _lib
from . import ops
from .geometry import Geometry
from .model import _Plan, _round_up
from .weights import ConvTcWeights, PreparedWeights

logger = logging.getLogger(__name__)


def _split(n: int, parts: int) -> List[int]:
    return [(r * n) // parts for r in range(parts + 1)]


class DomainLayout:
    """Band boundaries and unit ownership for every stage."""

    def __init__(self, geo: Geometry, world: int):
        self.world = world
        # the band plan exchanges ONE halo row for the stage 1-3 cross-embeds and builds nested bands from stride 2
        for st in geo.stages:
            for br in st.branches:
                if br.stride != 2:
                    raise NotImplementedError(f"stage {st.index}: cross-embed stride {br.stride}; the band layout needs 2")
                if st.index > 0 and (br.kernel - br.stride) // 2 > 1:
                    raise NotImplementedError(
                        f"stage {st.index}: cross-embed kernel {br.kernel} needs a {(br.kernel - br.stride) // 2}-row halo; "
                        "the decomposed plan exchanges one row for stages 1-3")
        rb3 = _split(geo.stages[3].h, world)
        self.rb = [[b * 2 ** (3 - s) for b in rb3] for s in range(4)]  # band row boundaries per stage
        self.units = []
        for st in geo.stages:
            ws, wg = st.local_window, st.global_window
            if st.h % (ws * wg) or st.w % (ws * wg):
                raise NotImplementedError(
                    f"stage {st.index}: grid {st.h}x{st.w} is not a multiple of local*global window {ws * wg}; "
                    "the unit decomposition needs that")
            a, b = st.h // (ws * wg), st.w // (ws * wg)
            self.units.append(dict(a=a, b=b, hu=ws * wg, wu=ws * wg, n=a * b, ub=_split(a * b, world)))
        if any(u["n"] < world for u in self.units):
            raise NotImplementedError("fewer attention units than ranks")
        if any(self.rb[3][r + 1] == self.rb[3][r] for r in range(world)):
            raise NotImplementedError("more ranks than stage-3 grid rows")


def _pixel_maps(st, lay: DomainLayout):
    """Per pixel of the stage grid: (band owner, band-local index, unit owner, unit-local index)."""
    s = st.index
    H, W = st.h, st.w
    ws = st.local_window
    u = lay.units[s]
    y = torch.arange(H)[:, None].expand(H, W)
    x = torch.arange(W)[None, :].expand(H, W)
    wy, s1 = y // ws, y % ws
    wx, s2 = x // ws, x % ws
    ua, i = wy % u["a"], wy // u["a"]
    ub_, j = wx % u["b"], wx // u["b"]
    uid = ua * u["b"] + ub_
    Y, X = i * ws + s1, j * ws + s2
    rb = torch.tensor(lay.rb[s])
    ubnd = torch.tensor(u["ub"])
    band_owner = torch.searchsorted(rb, y.contiguous(), right=True) - 1
    unit_owner = torch.searchsorted(ubnd, uid.contiguous(), right=True) - 1
    band_local = (y - rb[band_owner]) * W + x
    unit_local = ((uid - ubnd[unit_owner]) * u["hu"] + Y) * u["wu"] + X
    return band_owner.reshape(-1), band_local.reshape(-1), unit_owner.reshape(-1), unit_local.reshape(-1)


class _Comm:
    """Point-to-point plumbing inside the domain group (ranks are positions within the group)."""

    def __init__(self, rank: int, world: int, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.peers = [r if group is None else dist.get_global_rank(group, r) for r in range(world)]

    def recv(self, t, r):
        return dist.P2POp(dist.irecv, t, self.peers[r], self.group)

    def send(self, t, r):
        return dist.P2POp(dist.isend, t, self.peers[r], self.group)

    @staticmethod
    def run(reqs):
        if reqs:
            for w in dist.batch_isend_irecv(reqs):
                w.wait()

    def all_reduce(self, t):
        dist.all_reduce(t, group=self.group)

    def all_to_all(self, out, inp, out_splits, in_splits):
        dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits, group=self.group)


class Exchange:
    """band <-> unit re-layout of a stage's residual stream for this rank (index lists are computed once)."""

    def __init__(self, st, lay: DomainLayout, comm: Optional[_Comm], device, rank: Optional[int] = None):
        bo, bl, uo, ul = _pixel_maps(st, lay)
        world, rank = lay.world, (comm.rank if rank is None else rank)
        self.comm = comm  # None: the peer-memory path only (in-process ranks of the CPU tests)
        self.rank, self.world, self.d = rank, world, st.dim
        send_idx, recv_idx = [], []
        self.send_counts, self.recv_counts = [], []
        for r in range(world):  # band side: what I send to r, ordered by r's unit-local index
            m = (bo == rank) & (uo == r)
            order = torch.argsort(ul[m])
            send_idx.append(bl[m][order])
            self.send_counts.append(int(m.sum()))
        for q in range(world):  # unit side: what I receive from q (already in unit-local order)
            m = (bo == q) & (uo == rank)
            recv_idx.append(torch.sort(ul[m]).values)
            self.recv_counts.append(int(m.sum()))
        self.n_band = int((bo == rank).sum())
        self.n_unit = int((uo == rank).sum())
        cat_send = torch.cat(send_idx)
        cat_recv = torch.cat(recv_idx)
        inv_recv = torch.empty(self.n_unit, dtype=torch.long)
        inv_recv[cat_recv] = torch.arange(self.n_unit)
        inv_send = torch.empty(self.n_band, dtype=torch.long)
        inv_send[cat_send] = torch.arange(self.n_band)
        i32 = dict(device=device, dtype=torch.int32)
        self.send_idx = cat_send.to(**i32)    # band-local rows in send order
        self.recv_idx = cat_recv.to(**i32)    # unit-local rows in receive order
        self.inv_recv = inv_recv.to(**i32)    # unit-local position -> row of the receive buffer
        self.inv_send = inv_send.to(**i32)    # band-local position -> row of the (reverse) receive buffer
        self.send_off = [0]
        for c in self.send_counts:
            self.send_off.append(self.send_off[-1] + c)
        self.recv_off = [0]
        for c in self.recv_counts:
            self.recv_off.append(self.recv_off[-1] + c)
        # peer-memory path: every row is stored straight into its owner's buffer (no pack / unpack)
        mb, mu = bo == rank, uo == rank
        self.b2u = tuple(t[mb].contiguous().to(**i32) for t in (bl, uo, ul))   # my band rows -> (unit owner, unit-local row)
        self.u2b = tuple(t[mu].contiguous().to(**i32) for t in (ul, bo, bl))   # my unit rows -> (band owner, band-local row)

    def band_to_unit(self, band, ld_band, unit, sbuf, rbuf):
        d = self.d
        ops.gather_rows(band, ld_band, self.send_idx, sbuf, d, self.n_band, d)
        self.comm.all_to_all(rbuf[: self.n_unit * d], sbuf[: self.n_band * d], [c * d for c in self.recv_counts],
                             [c * d for c in self.send_counts])
        ops.gather_rows(rbuf, d, self.inv_recv, unit, d, self.n_unit, d)

    def band_to_unit_peer(self, band, ld_band, peer, xu_off, site):
        src_idx, dst_rank, dst_idx = self.b2u
        peer.scatter_rows(band, ld_band, src_idx, dst_rank, dst_idx, xu_off, self.d, self.n_band, self.d, site)
        peer.wait_all(site)

    def unit_to_band_peer(self, unit, peer, eb_off, ld_band, site):
        src_idx, dst_rank, dst_idx = self.u2b
        peer.scatter_rows(unit, self.d, src_idx, dst_rank, dst_idx, eb_off, ld_band, self.n_unit, self.d, site)
        peer.wait_all(site)

    def unit_to_band(self, unit, band, ld_band, sbuf, rbuf):
        d = self.d
        ops.gather_rows(unit, d, self.recv_idx, sbuf, d, self.n_unit, d)
        self.comm.all_to_all(rbuf[: self.n_band * d], sbuf[: self.n_unit * d], [c * d for c in self.send_counts],
                             [c * d for c in self.recv_counts])
        ops.gather_rows(rbuf, d, self.inv_send, band, ld_band, self.n_band, d)


def _halo_exchange(tensors, rows: int, comm: _Comm):
    """One-row halo exchange of band tensors shaped [rows + 2, W, C] (row 0 / rows+1 are the halos)."""
    rank, world = comm.rank, comm.world
    reqs = []
    for t in tensors:
        if rank > 0:
            reqs.append(comm.recv(t[0], rank - 1))
            reqs.append(comm.send(t[1], rank - 1))
        if rank < world - 1:
            reqs.append(comm.recv(t[rows + 1], rank + 1))
            reqs.append(comm.send(t[rows], rank + 1))
    comm.run(reqs)


def _shift_taps(w: ConvTcWeights, dy: int) -> ConvTcWeights:
    """Copy of the tensor-core conv weights whose tap rows are shifted by ``dy`` (input band carries a halo row)."""
    import copy
    import ctypes

    out = copy.copy(w)
    flat = [int(v) for v in w.taps_host]
    for k in range(0, len(flat), 2):
        flat[k] += dy
    out.taps_host = (ctypes.c_int32 * len(flat))(*flat)
    return out


class DomainPlan(_Plan):
    """Launch plan of one rank.  Input: the full state (replicated); output: the full prediction on every rank."""

    def __init__(self, geo: Geometry, wts: PreparedWeights, rank: int, world: int, device, group=None, peer=None):
        """``peer``: an already constructed peer communicator (the in-process stand-in of the CPU tests); by default a CUDA
        plan builds its own ``peer.PeerComm`` and a CPU plan uses the ``torch.distributed`` path."""
        self.wx = geo.variant == "wxformer"
        if wts.embed0_toep is None or wts.head_tc is None or any(c is None for brs in wts.embeds_tc[1:] for c in brs):
            raise NotImplementedError("domain decomposition needs the tensor-core path (channel counts % 4 == 0)")
        if self.wx and (wts.head2_tc is None or any(u.sharp_tc is None for u in wts.ups)):
            raise NotImplementedError("domain decomposition of the wxformer variant needs output channels % 8 == 0")
        self.geo, self.batch = geo, 1
        self.rank, self.world = rank, world
        self.comm = _Comm(rank, world, group) if not getattr(peer, "fake", False) else None
        self.tensor_cores = True
        self.attention_tc = True
        self.attn_simt_small = False
        self.toeplitz = True
        self.lay = lay = DomainLayout(geo, world)
        g = geo
        f32 = dict(device=device, dtype=torch.float32)
        f16 = dict(device=device, dtype=torch.float16)
        self.ld0 = 64
        self.xp = None
        self.xp_planes = (torch.empty((1, g.h_pad, g.w_pad, 64), **f16), torch.empty((1, g.h_pad, g.w_pad, 64), **f16))
        rows_of = lambda r: [lay.rb[s][r + 1] - lay.rb[s][r] for s in range(4)]  # noqa: E731
        nu_of = lambda r: [lay.units[s]["ub"][r + 1] - lay.units[s]["ub"][r] for s in range(4)]  # noqa: E731
        self.rows_all = [rows_of(r) for r in range(world)]
        self.rows = self.rows_all[rank]
        self.nu = nu_of(rank)
        m_unit = [self.nu[s] * lay.units[s]["hu"] * lay.units[s]["wu"] for s in range(4)]
        m_band = [self.rows[s] * g.stages[s].w for s in range(4)]
        big = _round_up(max(max(m_unit[s], m_band[s]) * g.stages[s].dim for s in range(4)), 8)
        self.ln = torch.empty(big, **f32)
        self.scratch = torch.empty(4 * big, **f32)
        self.ln16 = self.ln.view(torch.float16)
        self.scratch16 = self.scratch.view(torch.float16)
        self.sbuf = torch.empty(big, **f32)
        self.rbuf = torch.empty(big, **f32)
        h3 = self.halo3 = 1 if self.wx else 0

        # ---- buffers other ranks write into (halo rows, re-layout targets, GroupNorm sums): peer arena or plain tensors ----
        def shapes(r):
            rows, nu = rows_of(r), nu_of(r)
            sh = {}
            for s in range(4):
                st = g.stages[s]
                sh[f"xu{s}"] = ((max(nu[s], 1), lay.units[s]["hu"], lay.units[s]["wu"], st.dim), torch.float32)
                sh[f"eb{s}"] = ((rows[s], st.w, st.dim), torch.float32)
                if s < 3:
                    for pl in ("hi", "lo"):
                        sh[f"catp{s}.{pl}"] = ((rows[s] + 2, st.w, 2 * st.dim), torch.float16)
            s3 = g.stages[3]
            for pl in ("hi", "lo"):
                sh[f"x3p.{pl}"] = ((rows[3] + 2 * h3, s3.w, s3.dim), torch.float16)
            for k, up in enumerate(g.ups):
                ro, wo, c = 2 * rows[3 - k], 2 * up.w_in, up.c_out
                for nm in (("sp", "bp", "up") if self.wx else ("sp", "bp")):
                    for pl in ("hi", "lo"):
                        sh[f"dec{k}.{nm}.{pl}"] = ((ro + 2, wo, c), torch.float16)
            if self.wx:
                for pl in ("hi", "lo"):
                    sh[f"vp.{pl}"] = ((2 * rows[0] + 2, g.w_dec, g.output_channels), torch.float16)
            sh["y_dec"] = ((g.h_dec, g.w_dec, g.output_channels), torch.float32)
            for i in range(2 * len(g.ups)):
                sh[f"gn_slots{i}"] = ((world, g.dim[0] * 2), torch.float64)
            return sh

        own = shapes(rank)
        numel = lambda shp: int(torch.Size(shp).numel())  # noqa: E731
        self.peer = peer
        self._off = {}
        mode = os.environ.get("WXF_DOMAIN_COMM", "peer")
        if peer is not None:
            every = self._every = [shapes(r) for r in range(world)]
            biggest = {k: max((every[r][k][0] for r in range(world)), key=numel) for k in own}
        elif world > 1 and torch.device(device).type == "cuda" and mode == "peer":
            from .peer import PeerComm, _ITEMSIZE

            every = self._every = [shapes(r) for r in range(world)]
            biggest = {k: max((every[r][k][0] for r in range(world)), key=numel) for k in own}
            total = sum((numel(biggest[k]) * _ITEMSIZE[own[k][1]] + 255) // 256 * 256 + 256 for k in own) + 64 * 256
            self.peer = PeerComm(rank, world, group, total, device)
        buf = {}
        for k, (shp, dt) in own.items():
            if self.peer is not None:
                buf[k], self._off[k] = self.peer.buffer(biggest[k], shp, dt)
            else:
                buf[k] = torch.zeros(shp, device=device, dtype=dt)
        self._buf = buf
        self.xu = [buf[f"xu{s}"] for s in range(4)]
        self.eb = [buf[f"eb{s}"] for s in range(4)]
        # skip/concat planes in band layout with one halo row above and below (zero at the domain edges)
        self.catp = [(buf[f"catp{s}.hi"], buf[f"catp{s}.lo"]) for s in range(3)]
        # stage-3 output planes; only the wxformer decoder's first conv3x3 reads a halo row of them
        self.x3p = (buf["x3p.hi"], buf["x3p.lo"])
        self.dec = []
        for k, up in enumerate(g.ups):
            ro, wo, c = 2 * self.rows[3 - k], 2 * up.w_in, up.c_out
            self.dec.append(dict(
                short=torch.empty((ro, wo, c), **f32), a=torch.empty((ro, wo, c), **f32),
                sp=(buf[f"dec{k}.sp.hi"], buf[f"dec{k}.sp.lo"]), bp=(buf[f"dec{k}.bp.hi"], buf[f"dec{k}.bp.lo"])))
            if self.wx:  # PixelShuffle output u (fp32 + planes with halo rows): input of the `sharp` convolution
                self.dec[-1].update(u=torch.empty((ro, wo, c), **f32), up=(buf[f"dec{k}.up.hi"], buf[f"dec{k}.up.lo"]))
        if self.wx:  # up_block4 = conv3x3 -> PixelShuffle -> conv3x3: the shuffled tensor, band rows + halo
            self.vp = (buf["vp.hi"], buf["vp.lo"])
        self.gn_sums = torch.empty((1, g.dim[0], 2), device=device, dtype=torch.float64)
        self.gn_stats = torch.empty((1, g.dim[0], 2), **f32)
        gn_bytes = max(ops.groupnorm_scratch_bytes(1, 2 * self.rows[3 - k] * 2 * up.w_in, up.c_out)
                       for k, up in enumerate(g.ups))
        self.gn_scratch = torch.empty(gn_bytes // 4 + 4, **f32)
        self.y_dec = buf["y_dec"]  # full decoder output (own band rows + one halo row from each neighbour)
        self._gn_calls = 0
        self._halo_names = {}
        for k in own:
            if k.endswith(".hi"):
                self._halo_names[buf[k].data_ptr()] = k[:-3]
        self.ex = [Exchange(g.stages[s], lay, self.comm, device, rank) for s in range(4)]
        self.steps: List[tuple] = []
        self.bias_tiles: List[torch.Tensor] = []
        self._keep: List[object] = []
        self._row_ranges()
        self._build(wts)

    # ------------------------------------------------------------------------------------------------------------
    def _build(self, wts: PreparedWeights):
        g, lay, rank, world = self.geo, self.lay, self.rank, self.world
        add = self._add
        for st in g.stages:
            s, d = st.index, st.dim
            rows, r0 = self.rows[s], lay.rb[s][rank]
            eb = self.eb[s]
            # ---- cross-embed in band layout: output rows [r0, r0 + rows) of the stage grid ----
            for bi, br in enumerate(st.branches):
                if s == 0:  # Toeplitz kernel on the replicated padded input, output-row offset r0
                    desc = ops.make_toeplitz_desc(self.xp_planes[0], self.xp_planes[1], wts.embed0_toep[bi], eb, B=1,
                                                  Hi=g.h_pad, Wi=g.w_pad, lda=64, Ho=rows, Wo=st.w, ldc=d,
                                                  c_off=br.c_off, oy_off=r0)
                    add(ops.cross_embed_toeplitz_tc, (desc,), f"embed0.k{br.kernel}",
                        2.0 * rows * st.w * br.c_out * st.c_in * br.kernel * br.kernel)
                else:       # input: previous stage's band planes (upper half of its concat buffer) with halo rows
                    dp = g.stages[s - 1].dim
                    src_hi, src_lo = self.catp[s - 1][0][..., dp:], self.catp[s - 1][1][..., dp:]
                    self._conv_tc(src_hi, src_lo, _shift_taps(wts.embeds_tc[s][bi], 1), f"embed{s}.k{br.kernel}", B=1,
                                  Hi=self.rows[s - 1] + 2, Wi=g.stages[s - 1].w, lda=2 * dp, Ho=rows, Wo=st.w, out=eb,
                                  ldc=d, c_off=br.c_off)
            # ---- band -> unit layout, transformer on the rank's units (a batch of mini-images), unit -> band ----
            u = lay.units[s]
            ex, xu = self.ex[s], self.xu[s]
            if self.peer is not None:
                add(ex.band_to_unit_peer, (eb, d, self.peer, self._off[f"xu{s}"], self.peer.site()), f"exchange.s{s}", 0,
                    8.0 * rows * st.w * d)
            else:
                add(ex.band_to_unit, (eb, d, xu, self.sbuf, self.rbuf), f"exchange.s{s}", 0, 8.0 * rows * st.w * d)
            if self.nu[s] > 0:
                self._transformer(wts.blocks[s], st, self.nu[s], u["hu"], u["wu"], xu, d, None)
            if self.peer is not None:
                add(ex.unit_to_band_peer, (xu, self.peer, self._off[f"eb{s}"], d, self.peer.site()), f"exchange.s{s}", 0,
                    8.0 * rows * st.w * d)
            else:
                add(ex.unit_to_band, (xu, eb, d, self.sbuf, self.rbuf), f"exchange.s{s}", 0, 8.0 * rows * st.w * d)
            # ---- stage output as operand planes (band layout): skip connection + next cross-embed ----
            m = rows * st.w
            if s < 3:
                hi, lo = self.catp[s]
                add(ops.split_f16x2, (eb, d, hi[1:, :, d:], lo[1:, :, d:], 2 * d, m, d), "split", 0, 8.0 * m * d)
            else:
                add(ops.split_f16x2, (eb, d, self.x3p[0][self.halo3:], self.x3p[1][self.halo3:], d, m, d), "split", 0,
                    8.0 * m * d)
                if self.wx:
                    self._add_halo(self.x3p, rows)
            if 1 <= s + 1 <= 3 and s < 3:
                # the next stage's k=4 branch needs one halo row of this stage's output
                self._add_halo((hi, lo), rows)

        # ---- decoder in band layout ----
        dec_planes, dec_ld, dec_rows, dec_halo = self.x3p, g.stages[3].dim, self.rows[3], self.halo3
        for k, (up, uw, skip) in enumerate(zip(g.ups, wts.ups, (2, 1, 0))):
            bufs = self.dec[k]
            rin, ro, wo, c = self.rows[3 - k], 2 * self.rows[3 - k], 2 * up.w_in, up.c_out
            n = ro * wo * c
            sp_hi, sp_lo = bufs["sp"]
            bp_hi, bp_lo = bufs["bp"]
            in_hi, in_lo = (dec_planes[0][dec_halo:], dec_planes[1][dec_halo:]) if dec_halo else dec_planes
            if self.wx:
                # UpBlockPS (wxformer/crossformer.py:137-162): conv3x3 C -> 4 C_out as four sub-pixel phases (one halo row of
                # the low-resolution input, exchanged for the skip half by the stage and here for the decoder half), then
                # x = u + sharp(u) with a halo row of u
                if k > 0:
                    self._add_halo(dec_planes, rin)
                u_hi, u_lo = bufs["up"]
                self._conv_tc(dec_planes[0], dec_planes[1], _shift_taps(uw.up_tc, 1), "dec_up", B=1, Hi=rin + 2, Wi=up.w_in,
                              lda=dec_ld, Ho=rin, Wo=up.w_in, out=bufs["u"], ldc=c, out_hi=u_hi[1:], out_lo=u_lo[1:], ldh=c)
                self._add_halo((u_hi, u_lo), ro)
                self._conv_tc(u_hi, u_lo, _shift_taps(uw.sharp_tc, 1), "dec_conv3x3", B=1, Hi=ro + 2, Wi=wo, lda=c, Ho=ro,
                              Wo=wo, out=bufs["short"], ldc=c, res=bufs["u"], ldr=c, out_hi=sp_hi[1:], out_lo=sp_lo[1:], ldh=c)
            else:
                # ConvTranspose k2 s2: no halo; fp32 shortcut + planes (interior rows of the halo'd buffer)
                self._conv_tc(in_hi, in_lo, uw.up_tc, "dec_up", B=1, Hi=rin, Wi=up.w_in, lda=dec_ld, Ho=rin, Wo=up.w_in,
                              out=bufs["short"], ldc=c, out_hi=sp_hi[1:], out_lo=sp_lo[1:], ldh=c)
            self._add_halo((sp_hi, sp_lo), ro)
            self._conv_tc(sp_hi, sp_lo, _shift_taps(uw.convs_tc[0], 1), "dec_conv3x3", B=1, Hi=ro + 2, Wi=wo, lda=c,
                          Ho=ro, Wo=wo, out=bufs["a"], ldc=c)
            count = float(4 * up.h_in * up.w_in) * (c // up.groups)  # global pixels x channels per group
            site = self._gn_site()
            if site is not None:
                add(self._gn_put, (bufs["a"], c, ro * wo, c, up.groups, site), "groupnorm_silu", 0, 4.0 * n)
            add(self._groupnorm, (bufs["a"], c, uw.gn_w[0], uw.gn_b[0], None, 0, bp_hi[1:], bp_lo[1:], c, 0, ro * wo, c,
                                  up.groups, count, site), "groupnorm_silu", 0, 8.0 * n)
            self._add_halo((bp_hi, bp_lo), ro)
            self._conv_tc(bp_hi, bp_lo, _shift_taps(uw.convs_tc[1], 1), "dec_conv3x3", B=1, Hi=ro + 2, Wi=wo, lda=c,
                          Ho=ro, Wo=wo, out=bufs["a"], ldc=c)
            chi, clo = self.catp[skip]
            site = self._gn_site()
            if site is not None:
                add(self._gn_put, (bufs["a"], c, ro * wo, c, up.groups, site), "groupnorm_silu", 0, 4.0 * n)
            add(self._groupnorm, (bufs["a"], c, uw.gn_w[1], uw.gn_b[1], bufs["short"], c, chi[1:], clo[1:], 2 * c, 0,
                                  ro * wo, c, up.groups, count, site), "groupnorm_silu", 0, 12.0 * n)
            dec_planes, dec_ld, dec_rows, dec_halo = self.catp[skip], 2 * c, ro, 1
        # up_block4 (ConvT k4 s2 p1) reads one halo row of the full concat buffer
        st0 = g.stages[0]
        self._add_halo(self.catp[0], self.rows[0])
        y_band = self.y_dec[2 * lay.rb[0][rank]: 2 * lay.rb[0][rank + 1]]
        if self.wx:
            # up_block4 of the wxformer variant (wxformer/crossformer.py:813-830): conv3x3 -> PixelShuffle -> conv3x3
            co = g.output_channels
            rv = 2 * self.rows[0]
            self._conv_tc(self.catp[0][0], self.catp[0][1], _shift_taps(wts.head_tc, 1), "dec_head", B=1,
                          Hi=self.rows[0] + 2, Wi=st0.w, lda=2 * st0.dim, Ho=self.rows[0], Wo=st0.w, out_hi=self.vp[0][1:],
                          out_lo=self.vp[1][1:], ldh=co)
            self._add_halo(self.vp, rv)
            self._conv_tc(self.vp[0], self.vp[1], _shift_taps(wts.head2_tc, 1), "dec_head", B=1, Hi=rv + 2, Wi=g.w_dec, lda=co,
                          Ho=rv, Wo=g.w_dec, out=y_band, ldc=co)
        else:
            self._conv_tc(self.catp[0][0], self.catp[0][1], _shift_taps(wts.head_tc, 1), "dec_head", B=1,
                          Hi=self.rows[0] + 2, Wi=st0.w, lda=2 * st0.dim, Ho=self.rows[0], Wo=st0.w, out=y_band,
                          ldc=g.output_channels)
        if self.peer is not None:
            # first / last row of the rank's decoder band -> the neighbours' buffers (bilinear resize halo)
            P, me, off = self.peer, self.rank, self._off["y_dec"]
            rb = g.w_dec * g.output_channels * 4
            d_lo, d_hi = 2 * lay.rb[0][me], 2 * lay.rb[0][me + 1]
            site = P.site()
            segs, sigs, waits = [], [], []
            if me > 0:
                segs.append((self.y_dec.data_ptr() + d_lo * rb, P.arena.base[me - 1] + off + d_lo * rb, rb))
                sigs.append(P.sig(me - 1, site, 1))
                waits.append(P.sig(me, site, 0))
            if me < world - 1:
                segs.append((self.y_dec.data_ptr() + (d_hi - 1) * rb, P.arena.base[me + 1] + off + (d_hi - 1) * rb, rb))
                sigs.append(P.sig(me + 1, site, 0))
                waits.append(P.sig(me, site, 1))
            add(self._peer_halo, (segs, sigs, waits), "halo", 0, 0)
        else:
            add(self._ydec_halo, (), "halo", 0, 0)

    # ------------------------------------------------------------------------------------------------------------
    def _add_halo(self, tensors, rows):
        """One-row halo exchange of a band tensor pair [rows + 2, W, C] with the two neighbours (NCCL send / recv, or stores
        into the neighbours' arenas + arrival counters)."""
        if self.peer is None:
            self._add(_halo_exchange, (tensors, rows, self.comm), "halo", 0, 0)
            return
        P, me, world = self.peer, self.rank, self.world
        name = self._halo_names[tensors[0].data_ptr()]
        site = P.site()
        segs, sigs, waits = [], [], []
        for pl, t in zip(("hi", "lo"), tensors):
            key = f"{name}.{pl}"
            off = self._off[key]
            rb = t[0].numel() * t.element_size()
            if me > 0:      # my first interior row -> the bottom halo row of the rank above
                rows_prev = self._every[me - 1][key][0][0] - 2
                segs.append((t.data_ptr() + rb, P.arena.base[me - 1] + off + (rows_prev + 1) * rb, rb))
            if me < world - 1:  # my last interior row -> the top halo row of the rank below
                segs.append((t.data_ptr() + rows * rb, P.arena.base[me + 1] + off, rb))
        if me > 0:
            sigs.append(P.sig(me - 1, site, 1))
            waits.append(P.sig(me, site, 0))
        if me < world - 1:
            sigs.append(P.sig(me + 1, site, 0))
            waits.append(P.sig(me, site, 1))
        self._add(self._peer_halo, (segs, sigs, waits), "halo", 0, 0)

    def _peer_halo(self, segs, sigs, waits):
        self.peer.put(segs, sigs)
        self.peer.wait(waits)

    def _gn_site(self):
        """Arrival counters + per-rank slots of one GroupNorm call (its own block: a fast rank may already be at the next
        GroupNorm while a rank two bands away still reads this one's sums)."""
        if self.peer is None:
            return None
        i = self._gn_calls
        self._gn_calls += 1
        return (self.peer.site(), self._buf[f"gn_slots{i}"], self._off[f"gn_slots{i}"])

    def _gn_put(self, x, ldx, hw_local, C, G, site):
        """GroupNorm, first half (peer path): the band's sums -> slot ``rank`` of every rank's arena."""
        ops.groupnorm_sums(x, ldx, self.gn_sums, self.gn_scratch, 1, hw_local, C, G)
        P, (sg, slots, off) = self.peer, site
        nb = G * 2 * 8
        segs = [(self.gn_sums.data_ptr(), P.arena.base[r] + off + self.rank * slots.shape[1] * 8, nb) for r in range(self.world)]
        P.put(segs, [P.sig(r, sg, self.rank) for r in range(self.world)])

    def _groupnorm(self, x, ldx, gamma, beta, res, ldr, y_hi, y_lo, ldh, h_off, hw_local, C, G, count, site=None):
        """GroupNorm + SiLU with statistics over the whole (all-rank) image: local sums, all-reduce, apply."""
        if site is not None:
            P, (sg, slots, off) = self.peer, site
            P.wait_all(sg)
            # every rank adds the slots in rank order: bit-identical statistics everywhere
            assert slots.shape[1] == 2 * G
            P.sum_slots(slots, self.gn_sums, 2 * G)
        else:
            ops.groupnorm_sums(x, ldx, self.gn_sums, self.gn_scratch, 1, hw_local, C, G)
            self.comm.all_reduce(self.gn_sums)
        ops.groupnorm_stats_from_sums(self.gn_sums, self.gn_stats, 1, G, count)
        ops.groupnorm_apply_f16x2(x, ldx, self.gn_stats, gamma, beta, res, ldr, y_hi, y_lo, ldh, h_off, 1, hw_local, C, G)

    def _row_ranges(self):
        """Host-side bookkeeping of the sharded boundary passes (all ranks compute all ranks' ranges).

        pad_rows[r]  : padded-input rows rank r's stage-0 band reads (cross-embed kernel k, stride s, padding (k-s)//2)
        out_rows[r]  : output rows rank r writes = rows whose upper bilinear source row lies in r's band of the decoder
        src_rows[r]  : rows of the UNPADDED state that pad_rows[r] maps to (earth / mirror index map)
        """
        g, lay, world = self.geo, self.lay, self.world
        st0 = g.stages[0]
        kmax = max(br.kernel for br in st0.branches)
        stride = st0.branches[0].stride
        p = (kmax - stride) // 2
        pt, pb = g.padding.pad_lat if g.padding.activate else (0, 0)
        H = g.image_height
        self.pad_rows, self.out_rows, self.src_rows = [], [], []
        # bilinear source rows, the kernel's arithmetic: src = fma(scale, dst + 0.5, -0.5) in fp32, clamped at 0
        scale = float(torch.tensor(g.h_crop, dtype=torch.float32) / torch.tensor(g.h_out, dtype=torch.float32))
        o = torch.arange(g.h_out, dtype=torch.float64)
        src = (scale * (o + 0.5) - 0.5).float().clamp_min(0)
        y0 = src.floor().long().clamp_max(g.h_crop - 1)
        y1 = (y0 + 1).clamp_max(g.h_crop - 1)
        top = pt
        for r in range(world):
            r0, r1 = lay.rb[0][r], lay.rb[0][r + 1]
            a, b = max(0, stride * r0 - p), min(g.h_pad, stride * (r1 - 1) - p + kmax)
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
            d_lo, d_hi = 2 * r0, 2 * r1  # band of the decoder output (up_block4 doubles the stage-0 grid)
            own = ((y0 + top) >= d_lo) & ((y0 + top) < d_hi)
            idx = own.nonzero().flatten()
            if idx.numel() == 0:
                self.out_rows.append((0, 0))
                continue
            o_lo, o_hi = int(idx[0]), int(idx[-1]) + 1
            if o_hi - o_lo != idx.numel() or int((y1 + top)[o_lo:o_hi].max()) > d_hi or int((y0 + top)[o_lo:o_hi].min()) < d_lo - 1:
                raise NotImplementedError("bilinear resize reaches beyond one halo row of the decoder band")
            self.out_rows.append((o_lo, o_hi))
        covered = sum(b - a for a, b in self.out_rows)
        if covered != g.h_out:
            raise NotImplementedError("output rows are not partitioned by the decoder bands")

    def _pad(self, x):
        g = self.geo
        lat, lon, mode = ((g.padding.pad_lat, g.padding.pad_lon, g.padding.mode) if g.padding.activate
                          else ((0, 0), (0, 0), "earth"))
        if self.peer is not None:
            self.peer.advance()  # one step number per forward: arrival counters are compared with it
        a, b = self.pad_rows[self.rank]
        ops.pad_to_pixel_major_f16x2(x, lat, lon, mode, 64, self.xp_planes[0], self.xp_planes[1], rows=(a, b - a))

    def _ydec_halo(self):
        """First / last row of the rank's decoder band -> the neighbours' full-size buffers (bilinear resize halo)."""
        lay, me, c = self.lay, self.rank, self.comm
        d_lo, d_hi = 2 * lay.rb[0][me], 2 * lay.rb[0][me + 1]
        reqs = []
        if me > 0:
            reqs.append(c.recv(self.y_dec[d_lo - 1], me - 1))
            reqs.append(c.send(self.y_dec[d_lo], me - 1))
        if me < self.world - 1:
            reqs.append(c.recv(self.y_dec[d_hi], me + 1))
            reqs.append(c.send(self.y_dec[d_hi - 1], me + 1))
        c.run(reqs)

    def _unpad(self, out):
        g = self.geo
        pt, pl = (g.padding.pad_lat[0], g.padding.pad_lon[0]) if g.padding.activate else (0, 0)
        o_lo, o_hi = self.out_rows[self.rank]
        ops.unpad_resize_to_nchw(self.y_dec, g.output_channels, out, 1, g.output_channels, g.h_dec, g.w_dec, pt, pl,
                                 g.h_crop, g.w_crop, g.h_out, g.w_out, rows=(o_lo, o_hi - o_lo))

    def exchange_rows(self, t: torch.Tensor, n_ch: int, need):
        """Make rows ``need[r]`` of ``t[0, :n_ch, 0]`` ([B=1, C, 1, H, W]) valid on every rank r, given that each rank holds
        its own ``out_rows`` (point-to-point, compact [n_ch, rows, W] messages)."""
        me, c = self.rank, self.comm
        view = t[0, :n_ch, 0]
        recvs, reqs = [], []
        for peer in range(self.world):
            if peer == me:
                continue
            po_lo, po_hi = self.out_rows[peer]
            mo_lo, mo_hi = self.out_rows[me]
            a, b = max(need[me][0], po_lo), min(need[me][1], po_hi)      # rows I need that the peer owns
            if b > a:
                buf = torch.empty((n_ch, b - a, view.shape[-1]), device=t.device, dtype=t.dtype)
                recvs.append((buf, a, b))
                reqs.append(c.recv(buf, peer))
            a, b = max(need[peer][0], mo_lo), min(need[peer][1], mo_hi)  # rows the peer needs that I own
            if b > a:
                reqs.append(c.send(view[:, a:b].contiguous(), peer))
        c.run(reqs)
        for buf, a, b in recvs:
            view[:, a:b].copy_(buf)

    def run_band(self, x: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """One forward on this rank's share: reads rows ``src_rows[rank]`` of ``x`` (full-size tensor, other rows are not
        touched) and writes rows ``out_rows[rank]`` of the full-size prediction tensor."""
        g = self.geo
        self._pad(x)
        for fn, args, _tag, _fl, _by in self.steps:
            fn(*args)
        if out is None:
            out = torch.empty((1, g.base_output_channels, g.output_frames, g.h_out, g.w_out), device=x.device,
                              dtype=torch.float32)
        self._unpad(out)
        return out

    def run(self, x: torch.Tensor) -> torch.Tensor:
        """Full state in, full prediction out on every rank (what ``model(x)`` returns)."""
        g = self.geo
        out = self.run_band(x)
        flat = out.view(1, g.output_channels, 1, g.h_out, g.w_out)  # [C, T] -> C*T: rows are exchanged for every frame
        self.exchange_rows(flat, g.output_channels, [(0, g.h_out)] * self.world)
        return out


class DomainParallelManager:
    """Process groups of the decomposition: ``domain_parallel_size`` consecutive ranks share one forecast
    (same roles as the reference's manager, credit/domain_parallel/manager.py:22-129; no gradient groups here)."""

    def __init__(self, world_size: int = None, domain_parallel_size: int = None):
        world_size = dist.get_world_size() if world_size is None else world_size
        domain_parallel_size = world_size if domain_parallel_size is None else domain_parallel_size
        if world_size % domain_parallel_size:
            raise ValueError(f"world_size ({world_size}) must be divisible by domain_parallel_size ({domain_parallel_size})")
        self.world_size, self.domain_parallel_size = world_size, domain_parallel_size
        self.data_parallel_size = world_size // domain_parallel_size
        rank = dist.get_rank()
        self.domain_rank = rank % domain_parallel_size
        self.dp_rank = rank // domain_parallel_size
        self.domain_group = None
        if self.data_parallel_size > 1:
            for i in range(self.data_parallel_size):
                ranks = list(range(i * domain_parallel_size, (i + 1) * domain_parallel_size))
                grp = dist.new_group(ranks)
                if rank in ranks:
                    self.domain_group = grp

    @property
    def domain_world_size(self):
        return self.domain_parallel_size


def convert_to_domain_parallel(model, manager: DomainParallelManager = None):
    """Switch a CrossFormerB200 to the decomposed forward (the call the reference makes at
    credit/domain_parallel/convert.py:86).  Unlike the reference's sharded-tensor contract, ``model(x)`` keeps taking the
    full state on every rank of the domain group and returns the full prediction; the decomposition is internal."""
    manager = manager or DomainParallelManager()
    model._domain = manager
    model._plans = {}
    return model
