"""Generate the golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports ``/root/reference`` (read-only), builds ``credit.models.crossformer.CrossFormer`` through the
reference's own registry path ``credit.models.load_model(conf)``, loads the deterministic synthetic
state dict (``miles_credit_b200.synth``) with ``strict=True``, runs ``eval()`` + ``no_grad`` forward in
fp32 and stores input-independent facts and outputs under ``tests/golden/``.  The GPU box has no
``/root/reference``; tests there read only the files written here.

The one obstacle to importing the reference is ``credit/models/crossformer.py:10`` pulling
``credit.postblock.gen1`` -> xarray (absent): a stub module is placed in ``sys.modules`` first; the
stubbed class is never instantiated because ``post_conf.activate`` is False.
"""
import json
import os
import sys
import types

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

stub = types.ModuleType("credit.postblock.gen1")


class PostBlock(nn.Module):  # never instantiated (post_conf.activate=False)
    def __init__(self, *a, **k):
        super().__init__()


stub.PostBlock = PostBlock
sys.modules["credit.postblock.gen1"] = stub

from credit.models import load_model  # noqa: E402
from credit.boundary_padding import TensorPadding  # noqa: E402

from miles_credit_b200.geometry import build_geometry, state_spec, workload  # noqa: E402
from miles_credit_b200.synth import state_checksum, synthetic_input, synthetic_state_dict  # noqa: E402
from oracle import crossformer_oracle as oracle  # noqa: E402

CASES = {
    # name: (kwargs, batch)
    "unit": (workload("unit"), 1),
    "unit_mirror_f2": (dict(
        frames=2, output_frames=1, image_height=40, image_width=80, levels=2, channels=2, surface_channels=1,
        input_only_channels=1, output_only_channels=2, patch_width=1, patch_height=1,
        dim=[32, 64, 128, 256], depth=[1, 1, 1, 1], global_window_size=[10, 5, 2, 1], local_window_size=5,
        cross_embed_kernel_sizes=[[4, 8, 16, 32], [2, 4], [2, 4], [2, 4]], cross_embed_strides=[2, 2, 2, 2],
        interp=True, use_spectral_norm=False,
        padding_conf=dict(activate=True, mode="mirror", pad_lat=[20, 20], pad_lon=[40, 40]),
        post_conf={"activate": False}), 2),
    # registry key "wxformer": credit/models/wxformer/crossformer.py (PixelShuffle decoder, ZeroPad2d cross-embed)
    "unit_wxformer": (dict(workload("unit"), variant="wxformer", output_only_channels=4, depth=[1, 1, 1, 1]), 1),
}
TAPS = ["s0.embed", "s0.l0.short_attn", "s0.out", "s2.out", "s3.out", "up_block1", "up_block3", "up_block4"]


def reference_taps(model):
    """Forward hooks on the reference module tree that mirror the oracle's tap names."""
    got = {}

    def keep(name):
        return lambda mod, inp, out: got.__setitem__(name, out.detach().clone())

    for s, (cel, tr) in enumerate(model.layers):
        cel.register_forward_hook(keep(f"s{s}.embed"))
        tr.register_forward_hook(keep(f"s{s}.out"))
    # residual sum after the first short attention of stage 0: recompute from hook on the attention module
    att = model.layers[0][1].layers[0][0]
    att.register_forward_hook(lambda mod, inp, out: got.__setitem__("s0.l0.short_attn", (out + inp[0]).detach().clone()))
    for n in ("up_block1", "up_block2", "up_block3", "up_block4"):
        getattr(model, n).register_forward_hook(keep(n))
    return got


def main():
    torch.set_num_threads(8)
    summary = {}
    for name, (kwargs, batch) in CASES.items():
        geo = build_geometry(**kwargs)
        sd = synthetic_state_dict(geo, seed=1000)
        ref_kwargs = {k: v for k, v in kwargs.items() if k != "variant"}
        model = load_model({"model": dict(ref_kwargs, type=kwargs.get("variant", "crossformer"))})
        ref_sd = model.state_dict()
        assert list(sorted(ref_sd)) == list(sorted(sd)), "state-dict key mismatch vs reference"
        for k in ref_sd:
            assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), (k, ref_sd[k].shape, sd[k].shape)
        model.load_state_dict(sd, strict=True)
        model.eval()
        got = reference_taps(model)
        x = synthetic_input(geo, batch=batch, seed=1000)
        with torch.no_grad():
            y = model(x)
            taps = {}
            y_or = oracle.forward(x, sd, geo, taps)
        err = float((y - y_or).abs().max() / y.abs().max())
        tap_err = {t: float((got[t] - taps[t]).abs().max() / got[t].abs().max()) for t in TAPS}
        print(name, "out", tuple(y.shape), "absmax", float(y.abs().max()), "oracle rel-max err", err)
        print("   taps", {k: f"{v:.2e}" for k, v in tap_err.items()})
        # fp64 error bar of the reference itself
        with torch.no_grad():
            y64 = model.double()(x.double()).float()
        self_err = float((y - y64).abs().max() / y64.abs().max())
        print("   reference fp32 vs fp64 self error", self_err)
        torch.save({
            "kwargs": kwargs, "batch": batch, "seed": 1000, "state_checksum": state_checksum(sd),
            "y": y.clone(), "taps": {t: got[t] for t in TAPS}, "ref_fp32_vs_fp64": self_err,
            "keys": {k: list(v.shape) for k, v in ref_sd.items()},
        }, os.path.join(HERE, f"{name}.pt"))
        summary[name] = {"oracle_rel_max_err": err, "ref_fp32_vs_fp64": self_err, "absmax": float(y.abs().max())}

    # state-dict key/shape tables of the BASELINE configs (no weights: shapes only)
    keys = {}
    for wl, variant in (("wxformer_6h_025deg", "crossformer"), ("smoke_1deg", "crossformer"),
                        ("wxformer_6h_025deg", "wxformer")):
        kw = workload(wl)
        with torch.device("meta"):
            m = load_model({"model": dict(kw, type=variant)})
        name = wl if variant == "crossformer" else f"{wl}:{variant}"
        keys[name] = {k: list(v.shape) for k, v in m.state_dict().items()}
        spec = state_spec(build_geometry(**dict(kw, variant=variant)))
        assert {k: list(s) for k, (s, _) in spec.items()} == keys[name], name
    json.dump(keys, open(os.path.join(HERE, "state_keys.json"), "w"))

    # padding known answers from the reference's TensorPadding (earth + mirror), small grid
    torch.manual_seed(5)
    xp = torch.randn(1, 3, 2, 9, 16)
    pads = {}
    for mode, lat, lon in (("earth", (3, 4), (5, 2)), ("mirror", (2, 3), (4, 4)), ("earth", (4, 4), (0, 0)),
                           ("earth", (9, 9), (16, 16))):
        tp = TensorPadding(mode=mode, pad_lat=lat, pad_lon=lon)
        out = tp.pad(xp)
        assert torch.equal(tp.unpad(out), xp)
        assert torch.equal(oracle.pad_field(xp, mode, lat, lon), out), (mode, lat, lon)
        pads[f"{mode}_{lat[0]}_{lat[1]}_{lon[0]}_{lon[1]}"] = out
    torch.save({"x": xp, "padded": pads}, os.path.join(HERE, "padding.pt"))
    json.dump(summary, open(os.path.join(HERE, "summary.json"), "w"), indent=1)
    print("wrote golden fixtures to", HERE)


if __name__ == "__main__":
    main()
