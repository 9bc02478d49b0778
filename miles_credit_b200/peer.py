"""NVLink peer-memory plumbing of the domain decomposition (csrc/wxf_peer.cu): one cudaMalloc'd arena per rank, exported
with cudaIpc and opened by every peer of the domain group, plus the three exchange patterns the decomposed forward needs —
halo rows to the two neighbours, the band <-> unit re-layout of a stage's residual stream, and the GroupNorm sums — as
plain stores into the consumer's arena followed by a system-scope counter increment; consumers poll the counter.

``torch.distributed`` is used once, at construction, to exchange the 64-byte IPC handles; the per-step data path has no
collective library call.  Reference for what is exchanged: credit/domain_parallel/halo_exchange.py:56-67,
credit/domain_parallel/layers.py:507-518.
"""

from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import lib as _lib
from . import ops

_ALIGN = 256
_ITEMSIZE = {torch.float32: 4, torch.float16: 2, torch.float64: 8, torch.int32: 4, torch.uint8: 1}
_TYPESTR = {torch.float32: "<f4", torch.float16: "<f2", torch.float64: "<f8", torch.int32: "<i4", torch.uint8: "|u1"}


class _Raw:
    """``__cuda_array_interface__`` view of arena memory so torch can alias it without owning it."""

    def __init__(self, ptr: int, shape, dtype):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": _TYPESTR[dtype],
                                         "data": (int(ptr), False), "version": 2, "strides": None}


class PeerArena:
    """A bump allocator over one device allocation that every rank of the group lays out identically (callers pass the
    maximum size over ranks), so an offset is valid in every rank's arena."""

    def __init__(self, rank: int, world: int, group, nbytes: int, device):
        self.rank, self.world, self.device = rank, world, device
        self.nbytes = int(nbytes)
        self.used = 0
        L = _lib.load()
        with torch.cuda.device(device):
            p = ctypes.c_void_p()
            _lib.check(L.wxf_peer_alloc(ctypes.byref(p), self.nbytes), "wxf_peer_alloc")
            self.local = int(p.value)
            handle = (ctypes.c_ubyte * 64)()
            _lib.check(L.wxf_peer_export(ctypes.c_void_p(self.local), handle), "wxf_peer_export")
            handles: List[Optional[bytes]] = [None] * world
            dist.all_gather_object(handles, bytes(handle), group=group)
            self.base: List[int] = []
            for r in range(world):
                if r == rank:
                    self.base.append(self.local)
                    continue
                q = ctypes.c_void_p()
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(handles[r])
                _lib.check(L.wxf_peer_open(buf, ctypes.byref(q)), "wxf_peer_open")
                self.base.append(int(q.value))
        self._closed = False

    def take(self, nbytes: int) -> int:
        off = self.used
        self.used = (off + int(nbytes) + _ALIGN - 1) // _ALIGN * _ALIGN
        if self.used > self.nbytes:
            raise RuntimeError(f"peer arena exhausted: {self.used} > {self.nbytes} bytes")
        return off

    def tensor(self, off: int, shape, dtype) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        if n == 0:
            return torch.empty(tuple(shape), device=self.device, dtype=dtype)
        with torch.cuda.device(self.device):
            return torch.as_tensor(_Raw(self.local + off, shape, dtype), device=self.device)

    def close(self, group=None):
        """Collective teardown: every rank unmaps its peers' arenas, then (after a barrier) frees its own.  Without an
        explicit ``close`` the arena simply lives until the process exits: freeing an exported allocation while a peer still
        maps it is undefined, and a destructor cannot run a barrier."""
        if self._closed:
            return
        self._closed = True
        L = _lib.load()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for r, b in enumerate(self.base):
                if r != self.rank:
                    L.wxf_peer_close(ctypes.c_void_p(b))
            dist.barrier(group=group)
            L.wxf_peer_free(ctypes.c_void_p(self.local))


def _ptr_array(vals: Sequence[int]):
    return (ctypes.c_void_p * len(vals))(*[ctypes.c_void_p(int(v)) if v else ctypes.c_void_p(None) for v in vals])


class PeerComm:
    """The exchange patterns on top of a ``PeerArena``.  Every exchange site owns 16 counters (one per possible sender) in
    every rank's arena; ``advance()`` is called once per forward and counters are compared with that step number."""

    def __init__(self, rank: int, world: int, group, nbytes: int, device):
        self.rank, self.world = rank, world
        self.arena = PeerArena(rank, world, group, nbytes, device)
        self.device = device
        off = self.arena.take(64)
        self._epoch_off = off            # [0]: step number, [1]: CTA election counter of the put / scatter kernels
        self.epoch_ptr = self.arena.local + off
        self.done_ptr = self.arena.local + off + 4
        self._keep = []

    # -- allocation ----------------------------------------------------------------------------------------------
    def buffer(self, shape_max, shape_own, dtype):
        """Arena block sized for the largest rank (``shape_max``), viewed with this rank's shape.  Returns (tensor, offset)."""
        n = 1
        for s in shape_max:
            n *= int(s)
        off = self.arena.take(max(n, 1) * _ITEMSIZE[dtype])
        return self.arena.tensor(off, shape_own, dtype), off

    def site(self) -> int:
        """Offset of a fresh block of 16 arrival counters."""
        return self.arena.take(64)

    def sig(self, r: int, site: int, slot: int) -> int:
        return self.arena.base[r] + site + 4 * slot

    # -- per-step operations (kernels on the current stream) -----------------------------------------------------------
    def advance(self):
        _lib.check(_lib.load().wxf_peer_epoch_advance(ctypes.c_void_p(self.epoch_ptr), ops._stream()), "wxf_peer_epoch_advance")
        ops.LAUNCHES += 1

    def put(self, segs, signals):
        """segs: [(src_ptr, dst_ptr, nbytes)], signals: [counter address]; data first, then every counter += 1."""
        if not segs and not signals:
            return
        src = _ptr_array([s[0] for s in segs])
        dst = _ptr_array([s[1] for s in segs])
        nb = (ctypes.c_int64 * max(len(segs), 1))(*[int(s[2]) for s in segs])
        sg = _ptr_array(signals) if signals else _ptr_array([0])
        st = _lib.load().wxf_peer_put(src, dst, nb, len(segs), sg, len(signals), ctypes.c_void_p(self.done_ptr), ops._stream())
        _lib.check(st, "wxf_peer_put")
        ops.LAUNCHES += 1

    def wait(self, signals):
        if not signals:
            return
        st = _lib.load().wxf_peer_wait(_ptr_array(signals), len(signals), ctypes.c_void_p(self.epoch_ptr), ops._stream())
        _lib.check(st, "wxf_peer_wait")
        ops.LAUNCHES += 1

    def scatter_rows(self, src: torch.Tensor, ld_src: int, src_idx, dst_rank, dst_idx, dst_off: int, ld_dst: int, n: int, d: int,
                     site: int):
        """Row i of the exchange: ``src[src_idx[i]]`` -> rank ``dst_rank[i]``'s buffer at ``dst_off``, row ``dst_idx[i]``; then the
        counter ``site[rank]`` of every rank += 1."""
        bases = _ptr_array([b + dst_off for b in self.arena.base])
        sigs = _ptr_array([self.sig(r, site, self.rank) for r in range(self.world)])
        st = _lib.load().wxf_peer_scatter_rows(src.data_ptr(), ld_src, src_idx.data_ptr(), dst_rank.data_ptr(), dst_idx.data_ptr(),
                                               bases, sigs, self.world, ld_dst, n, d, ctypes.c_void_p(self.done_ptr), ops._stream())
        _lib.check(st, "wxf_peer_scatter_rows")
        ops.LAUNCHES += 1

    def wait_all(self, site: int):
        self.wait([self.sig(self.rank, site, r) for r in range(self.world)])

    def sum_slots(self, slots: torch.Tensor, sums: torch.Tensor, n: int):
        st = _lib.load().wxf_sum_rank_slots(slots.data_ptr(), sums.data_ptr(), self.world, n, ops._stream())
        _lib.check(st, "wxf_sum_rank_slots")
        ops.LAUNCHES += 1

    def close(self, group=None):
        self.arena.close(group)
