"""Golden vector for the noise-injection ensemble variant from the UNMODIFIED reference class
``credit.models.wxformer.crossformer_ensemble.CrossFormerWithNoise`` (registry key ``crossformer-ensemble``), run in the
build container only:

    python tests/golden/make_golden_ensemble.py

The reference draws its randomness with ``torch.randn`` inside ``forward`` (a latent vector and a per-pixel field for each
of the six injection sites).  To make the forward reproducible, ``torch.randn`` is wrapped for the duration of the forward by
a function that draws from a seeded CPU generator and RECORDS every tensor it hands out; the fixture stores the recorded
draws next to the input and the output, so the CUDA path can be fed the very same noise.  The module itself is untouched.
"""
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

from miles_credit_b200.geometry import build_geometry, workload  # noqa: E402
from miles_credit_b200.synth import synthetic_input, synthetic_state_dict  # noqa: E402


def main():
    for name, extra in (("unit_ensemble", dict(encoder_noise=True, correlated=False)),
                        ("unit_ensemble_correlated", dict(encoder_noise=False, correlated=True))):
        kw = dict(workload("unit"), output_only_channels=4, depth=[1, 1, 1, 1])
        conf = dict(kw, type="crossformer-ensemble", noise_latent_dim=16, encoder_noise_factor=0.05, decoder_noise_factor=0.275,
                    freeze=True, **extra)
        torch.manual_seed(1234)
        model = load_model({"model": dict(conf)})
        geo = build_geometry(**kw)
        base = synthetic_state_dict(geo, seed=77)
        own = model.state_dict()
        g = torch.Generator().manual_seed(4321)
        for k, v in own.items():            # the noise layers keep their default init, perturbed so every factor matters
            if k in base:
                continue
            if k.endswith("modulation"):
                own[k] = 1.0 + 0.2 * torch.randn(v.shape, generator=g)
            elif k.endswith("noise_transform.weight") or k.endswith("noise_transform.bias"):
                own[k] = v + 0.05 * torch.randn(v.shape, generator=g)
        own.update(base)
        model.load_state_dict(own, strict=True)
        model.eval()
        x = synthetic_input(geo, batch=2, seed=77)

        draws = []
        gen = torch.Generator().manual_seed(99)
        real_randn = torch.randn

        def recording_randn(*size, device=None, **k):
            shape = size[0] if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else size
            t = real_randn(*shape, generator=gen)
            draws.append(t.clone())
            return t

        torch.randn = recording_randn
        try:
            with torch.no_grad():
                y = model(x)
        finally:
            torch.randn = real_randn
        keys = {k: list(v.shape) for k, v in model.state_dict().items()}
        noise_sd = {k: v.clone() for k, v in model.state_dict().items() if k not in base}
        torch.save({"kwargs": {k: v for k, v in conf.items() if k != "type"}, "seed": 77, "x": x, "y": y, "draws": draws,
                    "keys": keys, "noise_state": noise_sd}, os.path.join(HERE, f"{name}.pt"))
        print(name, "y", tuple(y.shape), "draws", [tuple(d.shape) for d in draws], "|y|max", float(y.abs().max()))


if __name__ == "__main__":
    main()
