"""CPU: the launch plan, descriptors and weight re-layout, executed through the C-ABI *emulator*
(tests/abi_emulator.py) and compared with the reference's golden vectors.  This checks host logic only —
the CUDA kernels themselves are checked by the -m gpu tests."""
import os

import pytest
import torch

from miles_credit_b200 import lib as wlib
from miles_credit_b200 import model as wmodel
from miles_credit_b200 import ops
from miles_credit_b200.geometry import build_geometry
from miles_credit_b200.synth import synthetic_input, synthetic_state_dict
from miles_credit_b200.weights import prepare

from abi_emulator import EmulatedLib


@pytest.fixture
def emulated(monkeypatch):
    emu = EmulatedLib()
    monkeypatch.setattr(wlib, "_lib", emu)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(ops, "_req", lambda *a, **k: None)
    return emu


@pytest.mark.parametrize("tensor_cores", [False, True])
@pytest.mark.parametrize("case", ["unit", "unit_mirror_f2", "unit_wxformer"])
def test_plan_through_emulated_abi_matches_reference(golden_dir, emulated, case, tensor_cores):
    fx = torch.load(os.path.join(golden_dir, f"{case}.pt"), weights_only=False)
    geo = build_geometry(**fx["kwargs"])
    sd = synthetic_state_dict(geo, seed=fx["seed"])
    wts = prepare(sd, geo, wmodel._round_up(geo.input_channels, 4))
    plan = wmodel._Plan(geo, wts, fx["batch"], torch.device("cpu"), tensor_cores)
    x = synthetic_input(geo, batch=fx["batch"], seed=fx["seed"])
    y = plan.run(x)
    err = float((y - fx["y"]).abs().max() / fx["y"].abs().max())
    assert y.shape == fx["y"].shape
    print(case, "tensor_cores" if tensor_cores else "exact", "rel-max", err)
    assert err < (2e-5 if tensor_cores else 1e-5), err
    d0 = geo.stages[0].dim
    s0 = plan.cat[0][..., d0:].permute(0, 3, 1, 2)
    assert float((s0 - fx["taps"]["s0.out"]).abs().max() / fx["taps"]["s0.out"].abs().max()) < 1e-5
    # one-token long windows (global window 1) skip the attention launch on the tensor-core path: output = v
    n_attn = 2 * sum(geo.depth)
    if tensor_cores:
        n_attn -= sum(dep for dep, st in zip(geo.depth, geo.stages) if st.global_window == 1)
    assert emulated.calls.count("attention_tc" if tensor_cores else "attention") == n_attn
    assert (emulated.calls.count("gemm_tc") == 8 * sum(geo.depth)) == tensor_cores
    if tensor_cores and case == "unit_wxformer":  # 6 embeds + 3 x (ps, sharp, 2 convs); 12 output channels: fp32 head
        assert emulated.calls.count("conv_tc") == 6 + 12
    if tensor_cores:  # stage 1-3 cross-embed (6) + decoder (3 x 3) run as tensor-core convolutions
        assert emulated.calls.count("conv_tc") >= 15
        assert emulated.calls.count("toeplitz") == 4  # the four stage-0 cross-embed branches


def test_plan_tensor_core_head_vs_oracle(emulated):
    """Output channel count divisible by 4: the k4s2p1 transposed-conv head also runs as a tensor-core conv."""
    from miles_credit_b200.geometry import workload
    from oracle import crossformer_oracle as oracle

    kw = dict(workload("unit"), output_only_channels=4, depth=[1, 1, 1, 1])
    geo = build_geometry(**kw)
    sd = synthetic_state_dict(geo, seed=11)
    wts = prepare(sd, geo, wmodel._round_up(geo.input_channels, 4))
    assert wts.head_tc is not None
    plan = wmodel._Plan(geo, wts, 1, torch.device("cpu"), True)
    x = synthetic_input(geo, batch=1, seed=11)
    y = plan.run(x)
    with torch.no_grad():
        ref = oracle.forward(x, sd, geo)
    assert float((y - ref).abs().max() / ref.abs().max()) < 2e-5
    assert emulated.calls.count("conv_tc") == 6 + 9 + 1


def test_plan_wxformer_tensor_core_head_vs_oracle(emulated):
    """wxformer variant with 16 output channels: both up_block4 convolutions run as tensor-core convs."""
    from miles_credit_b200.geometry import workload
    from oracle import crossformer_oracle as oracle

    kw = dict(workload("unit"), variant="wxformer", output_only_channels=8, depth=[1, 1, 1, 1])
    geo = build_geometry(**kw)
    assert geo.output_channels == 16
    sd = synthetic_state_dict(geo, seed=12)
    wts = prepare(sd, geo, wmodel._round_up(geo.input_channels, 4))
    assert wts.head_tc is not None and wts.head2_tc is not None
    plan = wmodel._Plan(geo, wts, 1, torch.device("cpu"), True)
    x = synthetic_input(geo, batch=1, seed=12)
    y = plan.run(x)
    with torch.no_grad():
        ref = oracle.forward(x, sd, geo)
    assert float((y - ref).abs().max() / ref.abs().max()) < 2e-5


@pytest.mark.parametrize("out_only", [16, 15])
def test_plan_wxformer_wide_output_head_does_not_alias(emulated, out_only):
    """wxformer variant with output_channels > dim[0]/2 (every shipped `type: wxformer` config: 71 vs 128, 84 vs 32):
    up_block4's PixelShuffle output and the decoder output must not share scratch (the second conv3x3 reads a halo of
    the first while other CTAs write the second).  The emulator asserts disjoint byte ranges; 24 channels = tensor-core
    head, 23 = fp32 head."""
    from miles_credit_b200.geometry import workload
    from oracle import crossformer_oracle as oracle

    kw = dict(workload("unit"), variant="wxformer", output_only_channels=out_only, depth=[1, 1, 1, 1])
    geo = build_geometry(**kw)
    assert geo.output_channels > geo.dim[0] // 2
    sd = synthetic_state_dict(geo, seed=15)
    wts = prepare(sd, geo, wmodel._round_up(geo.input_channels, 4))
    assert (wts.head2_tc is not None) == (geo.output_channels % 8 == 0)
    plan = wmodel._Plan(geo, wts, 1, torch.device("cpu"), True)
    x = synthetic_input(geo, batch=1, seed=15)
    y = plan.run(x)
    with torch.no_grad():
        ref = oracle.forward(x, sd, geo)
    assert float((y - ref).abs().max() / ref.abs().max()) < 2e-5


@pytest.mark.parametrize("variant", ["batch2", "no_padding", "no_interp", "mirror_batch2_frames2"])
def test_plan_edge_configurations_vs_oracle(emulated, variant):
    """Host logic on configurations the golden fixtures do not cover: batch > 1, padding switched off
    (crossformer.py:598-599 skipped), interpolation switched off (:631-632: the output keeps the cropped size)."""
    from miles_credit_b200.geometry import workload
    from oracle import crossformer_oracle as oracle

    kw = dict(workload("unit"), depth=[1, 1, 1, 1], output_only_channels=4)
    batch = 1
    if variant == "batch2":
        batch = 2
    elif variant == "no_padding":  # 96 x 144 is what the padded unit grid is: windows still divide every stage
        kw.update(image_height=96, image_width=144, padding_conf=dict(activate=False))
    elif variant == "no_interp":
        kw.update(interp=False)
    elif variant == "mirror_batch2_frames2":
        kw.update(frames=2, padding_conf=dict(activate=True, mode="mirror", pad_lat=[25, 26], pad_lon=[24, 24]),
                  image_height=45)
        batch = 2
    geo = build_geometry(**kw)
    sd = synthetic_state_dict(geo, seed=13)
    wts = prepare(sd, geo, wmodel._round_up(geo.input_channels, 4))
    plan = wmodel._Plan(geo, wts, batch, torch.device("cpu"), True)
    x = synthetic_input(geo, batch=batch, seed=13)
    y = plan.run(x)
    with torch.no_grad():
        ref = oracle.forward(x, sd, geo)
    assert y.shape == ref.shape, (y.shape, ref.shape)
    err = float((y - ref).abs().max() / ref.abs().max())
    print(variant, tuple(y.shape), "rel-max", err)
    assert err < 2e-5, err


def test_plan_with_simt_small_window_attention_vs_oracle(emulated, monkeypatch):
    """WXF_ATTN_SIMT_SMALL=1 (round-2 candidate): dilated groups of <= 8 tokens go through the CUDA-core attention kernel
    (fp32 qkv in, planes out); everything else stays on the tensor-core kernel."""
    from miles_credit_b200.geometry import workload
    from oracle import crossformer_oracle as oracle

    monkeypatch.setenv("WXF_ATTN_SIMT_SMALL", "1")
    kw = dict(workload("unit"), output_only_channels=4)  # global windows 8, 4, 2, 1: stage 2 has L = 4
    geo = build_geometry(**kw)
    sd = synthetic_state_dict(geo, seed=15)
    wts = prepare(sd, geo, wmodel._round_up(geo.input_channels, 4))
    plan = wmodel._Plan(geo, wts, 1, torch.device("cpu"), True)
    x = synthetic_input(geo, batch=1, seed=15)
    y = plan.run(x)
    with torch.no_grad():
        ref = oracle.forward(x, sd, geo)
    assert float((y - ref).abs().max() / ref.abs().max()) < 2e-5
    n_small = sum(dep for dep, st in zip(geo.depth, geo.stages) if 1 < st.global_window ** 2 <= 8)
    assert n_small == geo.depth[2]
    assert emulated.calls.count("attention") == n_small  # the f16x2 entry point logs through the fp32 emulation


def test_rollout_state_update_with_forcing_vs_oracle(emulated):
    """rollout.Rollout (non-decomposed): y = model(x); prognostic channels of x come from y, the first n_dynamic input-only
    channels from the forcing tensor, the rest is carried (update_x, datasets/gen_2/channel_utils.py:253-291)."""
    from miles_credit_b200.geometry import workload
    from miles_credit_b200.rollout import Rollout
    from oracle import crossformer_oracle as oracle

    kw = dict(workload("unit"), depth=[1, 1, 1, 1], output_only_channels=4)
    geo = build_geometry(**kw)
    sd = synthetic_state_dict(geo, seed=16)
    wts = prepare(sd, geo, wmodel._round_up(geo.input_channels, 4))
    plan = wmodel._Plan(geo, wts, 1, torch.device("cpu"), True)

    class _Model:  # what Rollout needs from CrossFormerB200 when it is not decomposed
        geometry = geo
        _domain = None

        def __call__(self, x):
            return plan.run(x)

    ro = Rollout(_Model())
    n_prog = geo.channels * geo.levels + geo.surface_channels
    assert ro.n_prog == n_prog and not ro.sharded and ro.own_rows(None) == (0, geo.h_out)
    x = synthetic_input(geo, batch=1, seed=16)
    xo = x.clone()
    n_dyn = 1
    for step in range(2):
        frc = torch.randn(1, n_dyn, 1, geo.image_height, geo.image_width)
        y = ro.step(x, frc, n_dyn)
        with torch.no_grad():
            yo = oracle.forward(xo, sd, geo)
        assert float((y - yo).abs().max() / yo.abs().max()) < 5e-5
        nxt = xo.clone()
        nxt[:, :n_prog] = yo[:, :n_prog]
        nxt[:, n_prog: n_prog + n_dyn] = frc
        xo = nxt
        assert float((x - xo).abs().max() / xo.abs().max()) < 5e-5
        assert torch.equal(x[:, n_prog + n_dyn:], xo[:, n_prog + n_dyn:])  # static channels carried bit for bit
