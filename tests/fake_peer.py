"""In-process stand-in for ``miles_credit_b200.peer.PeerComm`` (test infrastructure): every "rank" of a decomposed plan lives
in this process, its arena is a numpy buffer, a put is a memmove and waits are no-ops — valid because the tests execute the
ranks' launch plans in LOCK-STEP (step i of every rank before step i + 1 of any), so a consumer step never runs before the
producer step of a peer.  Exercises the host logic of the peer-memory decompositions (offsets, halo rows, counters, index
lists) on the CPU through the C-ABI emulator."""
import ctypes

import numpy as np
import torch

_ITEM = {torch.float32: 4, torch.float16: 2, torch.float64: 8, torch.int32: 4}
_NP = {torch.float32: np.float32, torch.float16: np.float16, torch.float64: np.float64, torch.int32: np.int32}


class _Arena:
    def __init__(self, bufs, rank):
        self.bufs = bufs
        self.base = [b.ctypes.data for b in bufs]
        self.local = self.base[rank]


class FakePeer:
    fake = True

    def __init__(self, rank, world, bufs):
        self.rank, self.world = rank, world
        self.arena = _Arena(bufs, rank)
        self.used = 64
        self.puts = 0

    def _take(self, nbytes):
        off = self.used
        self.used = (off + int(nbytes) + 255) // 256 * 256
        assert self.used <= self.arena.bufs[self.rank].nbytes, "fake arena too small"
        return off

    def buffer(self, shape_max, shape_own, dtype):
        n = int(np.prod(shape_max)) if len(shape_max) else 1
        off = self._take(max(n, 1) * _ITEM[dtype])
        cnt = int(np.prod(shape_own))
        arr = self.arena.bufs[self.rank][off: off + cnt * _ITEM[dtype]].view(_NP[dtype]).reshape(shape_own)
        return torch.from_numpy(arr), off

    def site(self):
        return self._take(64)

    def sig(self, r, site, slot):
        return self.arena.base[r] + site + 4 * slot

    def advance(self):
        pass

    def put(self, segs, signals):
        for src, dst, n in segs:
            assert n % 16 == 0 and src % 16 == 0 and dst % 16 == 0, "16-byte segments (wxf_peer_put)"
            ctypes.memmove(dst, src, n)
        self.puts += 1

    def wait(self, signals):
        pass

    def wait_all(self, site):
        pass

    def sum_slots(self, slots, sums, n):
        sums.view(-1)[:n].copy_(slots.view(self.world, -1)[:, :n].sum(0))

    def scatter_rows(self, src, ld_src, src_idx, dst_rank, dst_idx, dst_off, ld_dst, n, d, site):
        rows = src.reshape(-1).as_strided((int(src_idx.max()) + 1 if n else 0, d), (ld_src, 1))
        for r in range(self.world):
            m = dst_rank == r
            if int(m.sum()) == 0:
                continue
            di = dst_idx[m].long()
            n_dst = int(di.max()) + 1
            dst = np.ctypeslib.as_array(ctypes.cast(self.arena.base[r] + dst_off, ctypes.POINTER(ctypes.c_float)),
                                        shape=((n_dst - 1) * ld_dst + d,))
            torch.from_numpy(dst).as_strided((n_dst, d), (ld_dst, 1))[di] = rows[src_idx[m].long()]


def make_fake_world(world, nbytes):
    bufs = [np.zeros(nbytes, dtype=np.uint8) for _ in range(world)]
    return [FakePeer(r, world, bufs) for r in range(world)]


def run_lockstep(plans, x):
    """One forward of every rank's plan in lock-step; returns the per-rank outputs of ``_unpad``."""
    for p in plans:
        p._pad(x)
    n = len(plans[0].steps)
    assert all(len(p.steps) == n for p in plans), "every rank must enqueue the same number of steps"
    for i in range(n):
        for p in plans:
            fn, args = p.steps[i][0], p.steps[i][1]
            fn(*args)
    outs = []
    for p in plans:
        g = p.geo
        shape = (1, *g.out_shape)
        out = torch.full(shape, float("nan"))
        p._unpad(out)
        outs.append(out)
    return outs
