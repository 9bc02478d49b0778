"""Per-process stand-in for ``miles_credit_b200.peer.PeerComm`` over gloo (test infrastructure): every rank is its own
process with its own arena (numpy), a put becomes a header + payload message to the destination rank, a wait receives
messages (from any source) until the arrival counters it polls have been bumped.  Unlike tests/fake_peer.py (all ranks in one
process, lock-step) the ranks run concurrently, so message order, early arrivals from a fast neighbour and the counter
protocol are exercised the way the NVLink path uses them."""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from fake_peer import _ITEM, _NP

_SPAN = 1 << 40  # "address" range of a remote arena: base[r] = (r + 1) * _SPAN


class _Arena:
    def __init__(self, buf, rank, world):
        self.buf = buf
        self.local = buf.ctypes.data
        self.base = [self.local if r == rank else (r + 1) * _SPAN for r in range(world)]


class GlooPeer:
    fake = True  # the plans skip their torch.distributed _Comm (the tests drive the steps themselves)

    def __init__(self, rank, world, nbytes):
        self.rank, self.world = rank, world
        self.arena = _Arena(np.zeros(nbytes, dtype=np.uint8), rank, world)
        self.used = 64
        self.puts = 0
        self._arrived = {}      # local signal offset -> arrivals so far
        self._consumed = {}     # local signal offset -> arrivals already waited for
        self._pending = []      # isend handles + the tensors they read

    # ---- arena bookkeeping (same as FakePeer) ------------------------------------------------------------------------
    def _take(self, nbytes):
        off = self.used
        self.used = (off + int(nbytes) + 255) // 256 * 256
        assert self.used <= self.arena.buf.nbytes, "arena too small"
        return off

    def buffer(self, shape_max, shape_own, dtype):
        n = int(np.prod(shape_max)) if len(shape_max) else 1
        off = self._take(max(n, 1) * _ITEM[dtype])
        cnt = int(np.prod(shape_own))
        arr = self.arena.buf[off: off + cnt * _ITEM[dtype]].view(_NP[dtype]).reshape(shape_own)
        return torch.from_numpy(arr), off

    def site(self):
        return self._take(64)

    def sig(self, r, site, slot):
        return self.arena.base[r] + site + 4 * slot

    def advance(self):
        pass

    def _owner(self, addr):
        if self.arena.local <= addr < self.arena.local + self.arena.buf.nbytes:
            return self.rank, addr - self.arena.local
        r = addr // _SPAN - 1
        assert 0 <= r < self.world and r != self.rank, hex(addr)
        return r, addr - (r + 1) * _SPAN

    # ---- data path -----------------------------------------------------------------------------------------------------
    def put(self, segs, signals):
        """segs: (local source address, destination address, bytes); signals: one arrival counter per destination rank."""
        self.puts += 1
        by_rank = {}
        for src, dst, n in segs:
            r, off = self._owner(dst)
            by_rank.setdefault(r, []).append((src, off, n))
        sig_of = {}
        for s in signals:
            r, off = self._owner(s)
            assert r not in sig_of, "one signal per destination and put"
            sig_of[r] = off
        assert set(by_rank) <= set(sig_of), "a destination without a signal"
        for r, soff in sig_of.items():
            parts = by_rank.get(r, [])
            if r == self.rank:
                for src, off, n in parts:
                    ctypes.memmove(self.arena.local + off, src, n)
                self._arrived[soff] = self._arrived.get(soff, 0) + 1
                continue
            assert len(parts) <= 8
            hdr = torch.zeros(20, dtype=torch.int64)
            hdr[0], hdr[1], hdr[2] = self.rank, soff, len(parts)
            chunks = []
            for k, (src, off, n) in enumerate(parts):
                hdr[3 + 2 * k], hdr[4 + 2 * k] = off, n
                chunks.append(torch.from_numpy(np.ctypeslib.as_array(ctypes.cast(src, ctypes.POINTER(ctypes.c_uint8)), shape=(n,)).copy()))
            payload = torch.cat(chunks) if chunks else torch.zeros(1, dtype=torch.uint8)
            self._pending.append((dist.isend(hdr, r), hdr))
            self._pending.append((dist.isend(payload, r), payload))

    def scatter_rows(self, src, ld_src, src_idx, dst_rank, dst_idx, dst_off, ld_dst, n, d, site):
        """Row i: src[src_idx[i]] -> rank dst_rank[i]'s buffer at dst_off, row dst_idx[i]; then counter site[me] of every rank += 1."""
        self.puts += 1
        rows = src.reshape(-1).as_strided((int(src_idx.max()) + 1 if n else 0, d), (ld_src, 1))
        soff = site + 4 * self.rank
        for r in range(self.world):
            m = dst_rank == r
            di = dst_idx[m].long()
            vals = rows[src_idx[m].long()].contiguous()
            if r == self.rank:
                if di.numel():
                    n_dst = int(di.max()) + 1
                    dst = torch.from_numpy(self.arena.buf[dst_off: dst_off + ((n_dst - 1) * ld_dst + d) * 4].view(np.float32))
                    dst.as_strided((n_dst, d), (ld_dst, 1))[di] = vals
                self._arrived[soff] = self._arrived.get(soff, 0) + 1
                continue
            hdr = torch.zeros(20, dtype=torch.int64)
            hdr[0], hdr[1], hdr[2] = self.rank, soff, -1                   # kind -1: indexed rows
            hdr[3], hdr[4], hdr[5], hdr[6] = di.numel(), dst_off, ld_dst, d
            idx = di.clone() if di.numel() else torch.zeros(1, dtype=torch.int64)
            payload = vals.reshape(-1).clone() if di.numel() else torch.zeros(1)
            for t in (hdr, idx, payload):
                self._pending.append((dist.isend(t, r), t))

    def _receive_rows(self, hdr, sender):
        cnt, dst_off, ld_dst, d = (int(v) for v in hdr[3:7])
        idx = torch.zeros(max(cnt, 1), dtype=torch.int64)
        dist.recv(idx, src=sender)
        payload = torch.zeros(max(cnt * d, 1))
        dist.recv(payload, src=sender)
        if cnt:
            n_dst = int(idx.max()) + 1
            dst = torch.from_numpy(self.arena.buf[dst_off: dst_off + ((n_dst - 1) * ld_dst + d) * 4].view(np.float32))
            dst.as_strided((n_dst, d), (ld_dst, 1))[idx] = payload.view(cnt, d)
        soff = int(hdr[1])
        self._arrived[soff] = self._arrived.get(soff, 0) + 1

    def _receive_one(self):
        hdr = torch.zeros(20, dtype=torch.int64)
        sender = dist.recv(hdr)                      # from any source
        if int(hdr[2]) < 0:
            return self._receive_rows(hdr, sender)
        n_parts = int(hdr[2])
        total = sum(int(hdr[4 + 2 * k]) for k in range(n_parts))
        payload = torch.zeros(max(total, 1), dtype=torch.uint8)
        dist.recv(payload, src=sender)
        pos = 0
        for k in range(n_parts):
            off, n = int(hdr[3 + 2 * k]), int(hdr[4 + 2 * k])
            self.arena.buf[off: off + n] = payload[pos: pos + n].numpy()
            pos += n
        soff = int(hdr[1])
        self._arrived[soff] = self._arrived.get(soff, 0) + 1

    def wait(self, signals):
        for s in signals:
            r, soff = self._owner(s)
            assert r == self.rank, "a rank waits on its own counters"
            while self._arrived.get(soff, 0) <= self._consumed.get(soff, 0):
                self._receive_one()
            self._consumed[soff] = self._consumed.get(soff, 0) + 1
        self._pending = [(h, t) for h, t in self._pending if not h.is_completed()]

    def wait_all(self, site):
        self.wait([self.sig(self.rank, site, r) for r in range(self.world)])

    def sum_slots(self, slots, sums, n):
        sums.view(-1)[:n].copy_(slots.view(self.world, -1)[:, :n].sum(0))

    def finish(self):
        for h, _t in self._pending:
            h.wait()
        self._pending = []
