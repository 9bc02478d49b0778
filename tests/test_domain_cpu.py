"""CPU (gloo, world_size 2 and 3): the lat-lon domain decomposition of the forecast step (miles_credit_b200/domain.py).

Every rank runs its share of the launch plan through the C-ABI emulator, exchanges band<->unit layouts, halos and
GroupNorm sums over gloo, and must reproduce the single-device oracle on the full grid.  Also checks the pure index
math (each pixel owned exactly once in both layouts; units closed under short and long attention)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from miles_credit_b200.geometry import build_geometry, workload

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _kwargs(name="unit"):
    if name == "unit8":
        # 128x64 padded grid, windows 2 / (4, 2, 1, 1): 32, 32, 32, 8 attention units and 8 stage-3 rows, so 8 ranks get
        # one unit and one row each at the coarsest stage; the 16-row input halos span more than one neighbour
        return dict(workload("unit"), image_height=101, image_width=48, levels=2, depth=[1, 1, 1, 1],
                    global_window_size=[4, 2, 1, 1], local_window_size=2, output_only_channels=2,
                    padding_conf=dict(activate=True, mode="earth", pad_lat=[13, 14], pad_lon=[8, 8]))
    if name == "unit_wx":  # PixelShuffle decoder (registry keys wxformer / wxformer_base), 16 output channels
        return dict(workload("unit"), variant="wxformer", output_only_channels=8, depth=[1, 1, 1, 1])
    if name == "unit_mirror":  # reflect-in-latitude padding: the rows a rank's padding pass reads fold back at the poles
        return dict(workload("unit"), output_only_channels=4,
                    padding_conf=dict(activate=True, mode="mirror", pad_lat=[25, 26], pad_lon=[24, 24]))
    return dict(workload("unit"), output_only_channels=4)


def _worker(rank, world, port, out_dir, dps, name="unit"):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    from abi_emulator import EmulatedLib
    from miles_credit_b200 import lib as wlib
    from miles_credit_b200 import model as wmodel
    from miles_credit_b200 import ops
    from miles_credit_b200.domain import DomainParallelManager, DomainPlan
    from miles_credit_b200.synth import synthetic_input, synthetic_state_dict
    from miles_credit_b200.weights import prepare

    torch.set_num_threads(2)
    wlib._lib = EmulatedLib()
    ops._stream = lambda: 0
    ops._req = lambda *a, **k: None
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        geo = build_geometry(**_kwargs(name))
        sd = synthetic_state_dict(geo, seed=21)
        wts = prepare(sd, geo, wmodel._round_up(geo.input_channels, 4))
        dm = DomainParallelManager(world, dps)
        plan = DomainPlan(geo, wts, dm.domain_rank, dm.domain_world_size, torch.device("cpu"), dm.domain_group)
        x = synthetic_input(geo, batch=1, seed=21)
        y = plan.run(x)
        torch.save(y, os.path.join(out_dir, f"y{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dps,name", [(2, 2, "unit"), (3, 3, "unit"), (4, 2, "unit"), (3, 3, "unit_wx")])
def test_domain_decomposition_matches_oracle(tmp_path, world, dps, name):
    from miles_credit_b200.synth import synthetic_input, synthetic_state_dict
    from oracle import crossformer_oracle as oracle

    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), dps, name), nprocs=world, join=True)
    geo = build_geometry(**_kwargs(name))
    sd = synthetic_state_dict(geo, seed=21)
    x = synthetic_input(geo, batch=1, seed=21)
    with torch.no_grad():
        ref = oracle.forward(x, sd, geo)
    ys = [torch.load(os.path.join(tmp_path, f"y{r}.pt")) for r in range(world)]
    for r, y in enumerate(ys):
        err = float((y - ref).abs().max() / ref.abs().max())
        print("rank", r, "rel-max", err)
        assert err < 2e-5, (r, err)
    for y in ys[1:]:  # every rank holds the same full prediction, bit for bit
        assert torch.equal(y, ys[0])


class _FakeModel:
    """What rollout.Rollout needs from CrossFormerB200, with the plan built on CPU memory (emulated ABI)."""

    def __init__(self, geo, plan, manager):
        self.geometry, self._domain, self._plans = geo, manager, {0: plan}

    def _plan_for(self, x):
        return x, self._plans[0]


def _rollout_worker(rank, world, port, out_dir, steps, name="unit"):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    from abi_emulator import EmulatedLib
    from miles_credit_b200 import lib as wlib
    from miles_credit_b200 import model as wmodel
    from miles_credit_b200 import ops
    from miles_credit_b200.domain import DomainParallelManager, DomainPlan
    from miles_credit_b200.rollout import Rollout
    from miles_credit_b200.synth import synthetic_input, synthetic_state_dict
    from miles_credit_b200.weights import prepare

    torch.set_num_threads(2)
    wlib._lib = EmulatedLib()
    ops._stream = lambda: 0
    ops._req = lambda *a, **k: None
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        geo = build_geometry(**_kwargs(name))
        sd = synthetic_state_dict(geo, seed=21)
        wts = prepare(sd, geo, wmodel._round_up(geo.input_channels, 4))
        dm = DomainParallelManager(world, world)
        plan = DomainPlan(geo, wts, dm.domain_rank, dm.domain_world_size, torch.device("cpu"), dm.domain_group)
        ro = Rollout(_FakeModel(geo, plan, dm))
        assert ro.sharded
        x = synthetic_input(geo, batch=1, seed=21)
        for _ in range(steps):
            y = ro.step(x)
        o_lo, o_hi = ro.own_rows(x)
        torch.save({"rows": (o_lo, o_hi), "y": y[..., o_lo:o_hi, :].clone(), "src": plan.src_rows[rank],
                    "x": x[..., plan.src_rows[rank][0]: plan.src_rows[rank][1], :].clone()},
                   os.path.join(out_dir, f"r{rank}.pt"))
        full = ro.gather(y.clone())
        torch.save(full, os.path.join(out_dir, f"full{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,name", [(2, "unit"), (3, "unit"), (8, "unit8"), (3, "unit_mirror")])
def test_sharded_rollout_matches_oracle_rollout(tmp_path, world, name):
    """Two steps with the state kept sharded (halo rows only) == two oracle steps on the full state."""
    from miles_credit_b200.synth import synthetic_input, synthetic_state_dict
    from oracle import crossformer_oracle as oracle

    steps = 2
    mp.spawn(_rollout_worker, args=(world, _free_port(), str(tmp_path), steps, name), nprocs=world, join=True)
    geo = build_geometry(**_kwargs(name))
    sd = synthetic_state_dict(geo, seed=21)
    x = synthetic_input(geo, batch=1, seed=21)
    n_prog = geo.channels * geo.levels + geo.surface_channels
    with torch.no_grad():
        for _ in range(steps):
            y = oracle.forward(x, sd, geo)
            x = x.clone()
            x[:, :n_prog] = y[:, :n_prog]  # update_x (datasets/gen_2/channel_utils.py:253-291)
    got = torch.zeros_like(y)
    covered = 0
    for r in range(world):
        d = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        a, b = d["rows"]
        got[..., a:b, :] = d["y"]
        covered += b - a
        sa, sb = d["src"]
        assert sa <= a and b <= sb  # a rank keeps at least its own rows of the state current
        ex = float((d["x"] - x[..., sa:sb, :]).abs().max() / x.abs().max())
        assert ex < 5e-5, (r, ex)
    assert covered == geo.h_out
    err = float((got - y).abs().max() / y.abs().max())
    print("sharded rollout rel-max", err)
    assert err < 5e-5, err
    for r in range(world):
        full = torch.load(os.path.join(tmp_path, f"full{r}.pt"))
        assert torch.equal(full, got)


@pytest.mark.parametrize("name,world", [("unit", 2), ("unit", 3), ("wxformer_6h_025deg", 2), ("wxformer_6h_025deg", 8)])
def test_layout_partitions(name, world):
    from miles_credit_b200.domain import DomainLayout, _pixel_maps

    geo = build_geometry(**workload(name))
    lay = DomainLayout(geo, world)
    for st in geo.stages:
        bo, bl, uo, ul = _pixel_maps(st, lay)
        n = st.h * st.w
        for owner, local in ((bo, bl), (uo, ul)):
            total = 0
            for r in range(world):
                loc = local[owner == r]
                assert loc.numel() == 0 or torch.equal(torch.sort(loc).values, torch.arange(loc.numel()))
                total += loc.numel()
            assert total == n
        # closure: the pixels of a short window, and of a dilated long group, share one unit
        u = lay.units[st.index]
        ws, wg = st.local_window, st.global_window
        y = torch.arange(st.h)[:, None].expand(st.h, st.w).reshape(-1)
        x = torch.arange(st.w)[None, :].expand(st.h, st.w).reshape(-1)
        gid = (uo * (1 << 20) + ul // (u["hu"] * u["wu"]))  # (rank, local unit) = global unit identity
        short_key = (y // ws) * st.w + (x // ws)
        nh, nw = st.h // wg, st.w // wg
        long_key = (y % nh) * st.w + (x % nw)  # long group (gh, gw) = (y mod H/gws, x mod W/gws), crossformer.py:279
        for key in (short_key, long_key):
            first = torch.zeros(int(key.max()) + 1, dtype=gid.dtype).scatter_(0, key, gid)
            assert torch.equal(first[key], gid)

class _NoComm:
    """DomainPlan._row_ranges needs geometry, layout and world only."""


@pytest.mark.parametrize("name,world", [("unit", 2), ("unit", 3), ("wxformer_6h_025deg", 2), ("wxformer_6h_025deg", 4),
                                        ("wxformer_6h_025deg", 8)])
def test_sharded_boundary_row_ranges(name, world):
    """Host bookkeeping of the sharded pad / un-pad: every rank pads the rows its stage-0 band reads, the output rows are a
    partition, a rank's bilinear resize stays within one halo row of its decoder band, and the state rows a rank keeps
    current contain the rows it writes (so update_x never reads a stale row)."""
    from miles_credit_b200.domain import DomainLayout, DomainPlan

    geo = build_geometry(**workload(name))
    plan = DomainPlan.__new__(DomainPlan)
    plan.geo, plan.world, plan.lay = geo, world, DomainLayout(geo, world)
    plan._row_ranges()
    st0 = geo.stages[0]
    kmax = max(br.kernel for br in st0.branches)
    p = (kmax - 2) // 2
    prev_hi = 0
    for r in range(world):
        r0, r1 = plan.lay.rb[0][r], plan.lay.rb[0][r + 1]
        a, b = plan.pad_rows[r]
        for oy in (r0, r1 - 1):  # first and last output row of the band: every tap row inside the image is padded
            lo, hi = 2 * oy - p, 2 * oy - p + kmax
            assert a <= max(lo, 0) and min(hi, geo.h_pad) <= b
        o_lo, o_hi = plan.out_rows[r]
        if o_hi > o_lo:
            assert o_lo == prev_hi
            prev_hi = o_hi
            s_lo, s_hi = plan.src_rows[r]
            assert s_lo <= o_lo and o_hi <= s_hi
    assert prev_hi == geo.h_out
    if name == "wxformer_6h_025deg" and world == 8:
        # 0.25 deg on 8 GPUs: 48-56 stage-0 rows per rank; at most 17 halo rows of state per side and neighbour
        for r in range(world):
            o_lo, o_hi = plan.out_rows[r]
            s_lo, s_hi = plan.src_rows[r]
            assert o_lo - s_lo <= 17 and s_hi - o_hi <= 17


@pytest.mark.parametrize("world,variant", [(2, "crossformer"), (3, "crossformer"), (2, "wxformer")])
def test_peer_memory_plan_in_lockstep_matches_oracle(world, variant, monkeypatch):
    """The NVLink peer-memory path of the decomposition (arena offsets, halo rows, per-row re-layout targets, GroupNorm slots)
    on in-process ranks: every rank's plan executed in lock-step through the C-ABI emulator, vs the CPU oracle."""
    from abi_emulator import EmulatedLib
    from fake_peer import make_fake_world, run_lockstep
    from miles_credit_b200 import lib as wlib
    from miles_credit_b200 import model as wmodel
    from miles_credit_b200 import ops
    from miles_credit_b200.domain import DomainPlan
    from miles_credit_b200.geometry import build_geometry, workload
    from miles_credit_b200.synth import synthetic_input, synthetic_state_dict
    from miles_credit_b200.weights import prepare
    from oracle import crossformer_oracle as oracle

    monkeypatch.setattr(wlib, "_lib", EmulatedLib())
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(ops, "_req", lambda *a, **k: None)
    kw = dict(workload("unit"), depth=[1, 1, 1, 1], output_only_channels=(8 if variant == "wxformer" else 4), variant=variant)
    geo = build_geometry(**kw)
    sd = synthetic_state_dict(geo, seed=41)
    wts = prepare(sd, geo, wmodel._round_up(geo.input_channels, 4))
    peers = make_fake_world(world, 96 << 20)
    plans = [DomainPlan(geo, wts, r, world, torch.device("cpu"), peer=peers[r]) for r in range(world)]
    x = synthetic_input(geo, batch=1, seed=41)
    outs = run_lockstep(plans, x)
    y = torch.full_like(outs[0], float("nan"))
    for r, o in enumerate(outs):
        lo, hi = plans[r].out_rows[r]
        y[..., lo:hi, :] = o[..., lo:hi, :]
    with torch.no_grad():
        ref = oracle.forward(x, sd, geo)
    err = float((y - ref).abs().max() / ref.abs().max())
    print(variant, "world", world, "peer path rel-max vs the oracle", err)
    assert torch.isfinite(y).all() and err < 2e-5, err
    assert all(p.puts > 0 for p in peers)


def _peer_gloo_worker(rank, world, port, out_dir, variant):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    from abi_emulator import EmulatedLib
    from gloo_peer import GlooPeer
    from miles_credit_b200 import lib as wlib
    from miles_credit_b200 import model as wmodel
    from miles_credit_b200 import ops
    from miles_credit_b200.domain import DomainPlan
    from miles_credit_b200.synth import synthetic_input, synthetic_state_dict
    from miles_credit_b200.weights import prepare

    torch.set_num_threads(2)
    wlib._lib = EmulatedLib()
    ops._stream = lambda: 0
    ops._req = lambda *a, **k: None
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        kw = dict(workload("unit"), depth=[1, 1, 1, 1], output_only_channels=(8 if variant == "wxformer" else 4), variant=variant)
        geo = build_geometry(**kw)
        sd = synthetic_state_dict(geo, seed=41)
        wts = prepare(sd, geo, wmodel._round_up(geo.input_channels, 4))
        peer = GlooPeer(rank, world, 96 << 20)
        plan = DomainPlan(geo, wts, rank, world, torch.device("cpu"), peer=peer)
        x = synthetic_input(geo, batch=1, seed=41)
        outs = []
        for _ in range(2):                            # the second forward reuses buffers, sites and counters
            plan._pad(x)
            for step in plan.steps:
                step[0](*step[1])
            out = torch.full((1, *geo.out_shape), float("nan"))
            plan._unpad(out)
            outs.append(out)
        peer.finish()
        torch.save({"out": outs, "rows": plan.out_rows[rank], "puts": peer.puts}, os.path.join(out_dir, f"p{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,variant", [(2, "crossformer"), (3, "wxformer")])
def test_peer_memory_plan_over_gloo_processes(tmp_path, world, variant):
    """The peer-memory path of the WXFormer decomposition with one process per rank over gloo (tests/gloo_peer.py: puts and
    indexed row scatters as messages, arrival counters, ranks running concurrently) vs the CPU oracle."""
    from miles_credit_b200.synth import synthetic_input, synthetic_state_dict
    from oracle import crossformer_oracle as oracle

    mp.spawn(_peer_gloo_worker, args=(world, _free_port(), str(tmp_path), variant), nprocs=world, join=True)
    kw = dict(workload("unit"), depth=[1, 1, 1, 1], output_only_channels=(8 if variant == "wxformer" else 4), variant=variant)
    geo = build_geometry(**kw)
    sd = synthetic_state_dict(geo, seed=41)
    x = synthetic_input(geo, batch=1, seed=41)
    with torch.no_grad():
        ref = oracle.forward(x, sd, geo)
    parts = [torch.load(os.path.join(tmp_path, f"p{r}.pt"), weights_only=False) for r in range(world)]
    for k in range(2):
        y = torch.full_like(ref, float("nan"))
        for p in parts:
            lo, hi = p["rows"]
            y[..., lo:hi, :] = p["out"][k][..., lo:hi, :]
        err = float((y - ref).abs().max() / ref.abs().max())
        print(variant, "world", world, "forward", k, "rel-max vs the oracle", err)
        assert torch.isfinite(y).all() and err < 2e-5, err
    assert all(p["puts"] > 0 for p in parts)
