"""CPU: the latitude-band decomposition of FuXi (miles_credit_b200/fuxi_domain.py) on 2 and 3 in-process ranks through the
C-ABI emulator and the lock-step peer stand-in (tests/fake_peer.py), against the golden output of the UNMODIFIED reference
module (tests/golden/unit_fuxi.pt): band geometry, halo rows, the 3-row exchanges of the shifted Swin blocks (incl. the
cyclic wrap and the row mask on the last rank), GroupNorm slots, the un-patchify halo."""
import os

import pytest
import torch

from miles_credit_b200 import fuxi as wfuxi
from miles_credit_b200 import lib as wlib
from miles_credit_b200 import ops
from miles_credit_b200.fuxi_domain import FuxiDomainPlan

from abi_emulator import EmulatedLib
from fake_peer import make_fake_world, run_lockstep


@pytest.fixture
def emulated(monkeypatch):
    emu = EmulatedLib()
    monkeypatch.setattr(wlib, "_lib", emu)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(ops, "_req", lambda *a, **k: None)
    return emu


@pytest.mark.parametrize("world", [1, 2, 3])
def test_fuxi_band_decomposition_matches_reference(golden_dir, emulated, world):
    fx = torch.load(os.path.join(golden_dir, "unit_fuxi.pt"), weights_only=False)
    geo = wfuxi.build_fuxi_geometry(**fx["kwargs"])
    assert geo.gh // geo.ws[0] == 3 and geo.shift == (1, 1)       # 3 window rows, shifted blocks wrap one row
    wts = wfuxi.prepare_fuxi(fx["state_dict"], geo)
    peers = make_fake_world(world, 64 << 20)
    plans = [FuxiDomainPlan(geo, wts, r, world, torch.device("cpu"), peer=peers[r]) for r in range(world)]
    # bands partition the token rows and the output rows
    assert plans[0].t[0][0] == 0 and plans[0].t[-1][1] == geo.th
    assert all(plans[0].t[r][1] == plans[0].t[r + 1][0] for r in range(world - 1))
    assert sum(b - a for a, b in plans[0].out_rows) == geo.h_out
    for b in range(fx["x"].shape[0]):                               # the decomposed plan takes one state at a time
        x = fx["x"][b: b + 1].contiguous()
        outs = run_lockstep(plans, x)
        y = torch.full_like(outs[0], float("nan"))
        for r, o in enumerate(outs):
            lo, hi = plans[r].out_rows[r]
            y[..., lo:hi, :] = o[..., lo:hi, :]
        assert torch.isfinite(y).all()
        ref = fx["y"][b: b + 1]
        err = float((y - ref).abs().max() / ref.abs().max())
        print(f"world {world}, sample {b}: rel-max vs the reference {err:.2e}")
        assert err < 2e-5, err
    if world > 1:
        n_shifted = sum(1 for i in range(geo.depth) if any(geo.block_shift(i)))
        tags = [s[2] for s in plans[0].steps]
        assert tags.count("shift_exchange") == 2 * n_shifted and tags.count("halo") == 6


def _gloo_worker(rank, world, port, out_dir, golden_dir):
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    import torch.distributed as dist
    from gloo_peer import GlooPeer

    torch.set_num_threads(2)
    wlib._lib = EmulatedLib()
    ops._stream = lambda: 0
    ops._req = lambda *a, **k: None
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        fx = torch.load(os.path.join(golden_dir, "unit_fuxi.pt"), weights_only=False)
        geo = wfuxi.build_fuxi_geometry(**fx["kwargs"])
        wts = wfuxi.prepare_fuxi(fx["state_dict"], geo)
        peer = GlooPeer(rank, world, 64 << 20)
        plan = FuxiDomainPlan(geo, wts, rank, world, torch.device("cpu"), peer=peer)
        outs = []
        for b in range(fx["x"].shape[0]):            # two forwards through the same plan: counters and buffers are reused
            x = fx["x"][b: b + 1].contiguous()
            plan._pad(x)
            for fn, args, _tag, _fl, _by in plan.steps:
                fn(*args)
            out = torch.full((1, *geo.out_shape), float("nan"))
            plan._unpad(out)
            outs.append(out)
        peer.finish()
        torch.save({"out": outs, "rows": plan.out_rows[rank], "puts": peer.puts}, os.path.join(out_dir, f"r{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_fuxi_band_decomposition_over_gloo_processes(golden_dir, tmp_path, world):
    """The same decomposition with every rank in its OWN process (gloo, world_size 2 and 3): puts are messages, waits poll
    arrival counters, ranks run concurrently (tests/gloo_peer.py) - vs the golden output of the unmodified reference module."""
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_gloo_worker, args=(world, port, str(tmp_path), golden_dir), nprocs=world, join=True)
    fx = torch.load(os.path.join(golden_dir, "unit_fuxi.pt"), weights_only=False)
    parts = [torch.load(os.path.join(tmp_path, f"r{r}.pt"), weights_only=False) for r in range(world)]
    assert all(p["puts"] > 0 for p in parts)
    for b in range(fx["x"].shape[0]):
        y = torch.full_like(parts[0]["out"][b], float("nan"))
        for p in parts:
            lo, hi = p["rows"]
            y[..., lo:hi, :] = p["out"][b][..., lo:hi, :]
        ref = fx["y"][b: b + 1]
        err = float((y - ref).abs().max() / ref.abs().max())
        print(f"gloo world {world}, sample {b}: rel-max vs the reference {err:.2e}")
        assert torch.isfinite(y).all() and err < 2e-5, err
