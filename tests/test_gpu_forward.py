"""GPU parity of the whole forecast step against the reference's golden vectors and the oracle."""
import os

import pytest
import torch

from miles_credit_b200.geometry import build_geometry, workload
from miles_credit_b200.model import CrossFormerB200
from miles_credit_b200.synth import synthetic_input, synthetic_state_dict
from oracle import crossformer_oracle as oracle

pytestmark = pytest.mark.gpu
TOL = 1e-4  # north-star budget: rel-max vs the reference fp32 forward


def relmax(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("exact", [False, True], ids=["tensorcore", "exactfp32"])
@pytest.mark.parametrize("case", ["unit", "unit_mirror_f2", "unit_wxformer"])
def test_forward_matches_reference_golden(golden_dir, case, exact):
    fx = torch.load(os.path.join(golden_dir, f"{case}.pt"), weights_only=False)
    geo = build_geometry(**fx["kwargs"])
    model = CrossFormerB200(**fx["kwargs"])
    model.exact_fp32 = exact
    model.load_state_dict(synthetic_state_dict(geo, seed=fx["seed"]), strict=True)
    model = model.cuda().eval()
    x = synthetic_input(geo, batch=fx["batch"], seed=fx["seed"])
    y = model(x.cuda())
    assert y.shape == fx["y"].shape and y.dtype == torch.float32
    err = relmax(y.cpu(), fx["y"])
    print(f"{case}: rel-max vs reference = {err:.3e} (reference fp32-vs-fp64 {fx['ref_fp32_vs_fp64']:.1e})")
    assert err < TOL
    # encoder stage outputs kept in the plan's skip buffers
    plan = next(iter(model._plans.values()))
    for s, name in ((0, "s0.out"), (2, "s2.out")):
        d = geo.stages[s].dim
        got = plan.cat[s][..., d:].permute(0, 3, 1, 2).cpu()
        assert relmax(got, fx["taps"][name]) < TOL, name
    assert relmax(plan.x3.permute(0, 3, 1, 2).cpu(), fx["taps"]["s3.out"]) < TOL


def test_forward_smoke_tiny_vs_oracle():
    """BASELINE config[0] (credit_smoke_test_v2.yml at 64x128): CUDA path vs the CPU oracle, same weights/input."""
    kw = workload("smoke_tiny")
    geo = build_geometry(**kw)
    sd = synthetic_state_dict(geo, seed=1000)
    x = synthetic_input(geo, batch=1, seed=1000)
    with torch.no_grad():
        ref = oracle.forward(x, sd, geo)
    model = CrossFormerB200(**kw)
    model.load_state_dict(sd, strict=True)
    y = model.cuda().eval()(x.cuda())
    err = relmax(y.cpu(), ref)
    print(f"smoke_tiny: rel-max vs oracle = {err:.3e}")
    assert err < TOL
    # determinism: a second call gives identical bits
    y2 = model(x.cuda())
    assert torch.equal(y, y2)


def test_rollout_two_steps_vs_oracle():
    kw = workload("unit")
    geo = build_geometry(**kw)
    sd = synthetic_state_dict(geo, seed=3)
    x = synthetic_input(geo, batch=1, seed=3)
    model = CrossFormerB200(**kw)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    from miles_credit_b200.rollout import Rollout

    ro = Rollout(model)
    xs = x.cuda().clone()
    xo = x.clone()
    for _ in range(2):
        y = ro.step(xs)
        with torch.no_grad():
            yo = oracle.forward(xo, sd, geo)
        assert relmax(y.cpu(), yo) < TOL
        xo = oracle.rollout_update(xo, yo, geo)
    assert relmax(xs.cpu(), xo) < TOL


def test_rollout_cuda_graph_replay_equals_eager_launches():
    """Rollout(graph=True) captures one step (kernels + state update) and replays it: same bits as eager launches."""
    kw = workload("unit")
    geo = build_geometry(**kw)
    sd = synthetic_state_dict(geo, seed=5)
    model = CrossFormerB200(**kw)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    from miles_credit_b200.rollout import Rollout

    x0 = synthetic_input(geo, batch=1, seed=5).cuda()
    n_dyn = max(geo.input_only_channels // 2, 1)
    frc = torch.randn(1, n_dyn, 1, geo.image_height, geo.image_width, device="cuda")
    xe, xg = x0.clone(), x0.clone()
    eager, graph = Rollout(model), Rollout(model, graph=True)
    for _ in range(3):
        ye = eager.step(xe, frc, n_dyn)
        yg = graph.step(xg, frc, n_dyn)
        assert torch.equal(ye, yg)
    assert torch.equal(xe, xg)
    assert graph.launches_per_replay > 0


@pytest.mark.parametrize("exact", [False, True], ids=["tensorcore", "exactfp32"])
def test_forward_wxformer_6h_1deg_vs_oracle(exact):
    """BASELINE config[1]: the 0.25-degree architecture (dims 128..1024, depth 2/2/8/2, lws 10) on the 181x360 grid."""
    kw = workload("wxformer_6h_1deg")
    geo = build_geometry(**kw)
    sd = synthetic_state_dict(geo, seed=1000, sn_iters=5)
    x = synthetic_input(geo, batch=1, seed=1000)
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    with torch.no_grad():
        ref = oracle.forward(x, sd, geo)
    model = CrossFormerB200(**kw)
    model.exact_fp32 = exact
    model.load_state_dict(sd, strict=True)
    y = model.cuda().eval()(x.cuda())
    err = relmax(y.cpu(), ref)
    print(f"wxformer_6h_1deg ({'exact fp32' if exact else 'tensor cores'}): rel-max vs oracle = {err:.3e}")
    assert err < TOL


def test_forward_wxformer_variant_tc_head_vs_oracle():
    """PixelShuffle decoder with a channel count that lets both up_block4 convs run on the tensor cores."""
    kw = dict(workload("unit"), variant="wxformer", output_only_channels=8)
    geo = build_geometry(**kw)
    sd = synthetic_state_dict(geo, seed=12)
    x = synthetic_input(geo, batch=2, seed=12)
    with torch.no_grad():
        ref = oracle.forward(x, sd, geo)
    model = CrossFormerB200(**kw)
    model.load_state_dict(sd, strict=True)
    y = model.cuda().eval()(x.cuda())
    err = relmax(y.cpu(), ref)
    print(f"wxformer variant (tensor-core head): rel-max vs oracle = {err:.3e}")
    assert err < TOL


@pytest.mark.parametrize("out_only", [16, 15], ids=["tc_head_24ch", "fp32_head_23ch"])
def test_forward_wxformer_variant_wide_output_vs_oracle(out_only):
    """output_channels > dim[0]/2, as in every shipped `type: wxformer` config (71 vs 128, 84 vs 32): up_block4's
    PixelShuffle output and the decoder output used to share scratch (a cross-CTA read/write race on the GPU)."""
    kw = dict(workload("unit"), variant="wxformer", output_only_channels=out_only)
    geo = build_geometry(**kw)
    assert geo.output_channels > geo.dim[0] // 2
    sd = synthetic_state_dict(geo, seed=15)
    x = synthetic_input(geo, batch=2, seed=15)
    with torch.no_grad():
        ref = oracle.forward(x, sd, geo)
    model = CrossFormerB200(**kw)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    y = model(x.cuda())
    err = relmax(y.cpu(), ref)
    print(f"wxformer variant, {geo.output_channels} output channels: rel-max vs oracle = {err:.3e}")
    assert err < TOL
    assert torch.equal(y, model(x.cuda()))  # a race would also show as run-to-run differences


def test_forward_full_grid_025deg_vs_reference():
    """BASELINE config[2] at FULL size: WXFormer-6h on the 721x1440 grid, CUDA path vs the UNMODIFIED reference module
    (oracle/_ref; the oracle restatement when the reference is not staged).  ~15-40 s of host time for the reference."""
    from oracle import ref_loader

    kw = workload("wxformer_6h_025deg")
    geo = build_geometry(**kw)
    sd = synthetic_state_dict(geo, seed=1000, sn_iters=5)
    x = synthetic_input(geo, batch=1, seed=1000)
    torch.set_num_threads(os.cpu_count() or 8)
    with torch.no_grad():
        if ref_loader.available():
            ref, who = ref_loader.reference_model(kw, sd)(x), "unmodified reference"
        else:
            ref, who = oracle.forward(x, sd, geo), "oracle"
    model = CrossFormerB200(**kw)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    y = model(x.cuda())
    assert y.shape == (1, 64, 1, 721, 1440)
    err = relmax(y.cpu(), ref)
    print(f"wxformer_6h_025deg full grid: rel-max vs {who} = {err:.3e}")
    assert err < TOL
    # size-independent properties at full size: bit-determinism, and batch independence (a batch of two states gives the
    # two single-state predictions up to the reduction order of the GroupNorm statistics: no cross-sample leakage through
    # the window packing)
    assert torch.equal(y, model(x.cuda()))
    x2 = torch.cat([x, torch.flip(x, dims=[-1])]).cuda()
    y2 = model(x2)
    assert relmax(y2[:1], y) < 1e-6
    assert relmax(y2[1:], model(x2[1:].contiguous())) < 1e-6
