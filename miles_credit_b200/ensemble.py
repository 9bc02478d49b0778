"""``CrossFormerWithNoiseB200`` — the noise-injection ensemble variant of the forecast step, drop-in for
``credit.models.wxformer.crossformer_ensemble.CrossFormerWithNoise`` (registry keys ``crossformer-ensemble`` /
``crossformer-style``, credit/models/__init__.py:51-60).

Same encoder / decoder kernels as ``CrossFormerB200``; at six sites — after the transformer stack of stages 0-2 (the tensor
that is both the skip connection and the next stage's input, crossformer_ensemble.py:139-146) and after UpBlock 1-3 (before
the skip concat, :148-162) — the feature map becomes ``x + noise_factor * eps * Linear(latent)[b, c] * modulation[c]``
(``StochasticDecompositionLayer.forward``, stochastic_decomposition_layer.py:21-42) in one extra pass that also writes the
operand planes of the consumer.  ``eps`` and ``latent`` come from a Philox generator inside the kernels, keyed by a seed and
a device-resident step counter, so a replayed CUDA graph draws fresh noise every step and ensemble members differ by seed.
``correlated=True`` draws one latent per forward for all sites (:121-122).

State-dict keys = the reference's: the CrossFormer keys plus, per site, ``{encoder_noise_layers.k | noise_injectN}.
{noise_transform.weight, noise_transform.bias, modulation, noise_factor}`` (the noise layers are created after
``apply_spectral_norm`` ran, so their Linear carries no spectral-norm hook, crossformer_ensemble.py:43-107).
"""

from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn

from . import ops
from .model import CrossFormerB200, _Holder, _Plan

SITES = ("encoder_noise_layers.0", "encoder_noise_layers.1", "encoder_noise_layers.2", "noise_inject1", "noise_inject2",
         "noise_inject3")


class NoiseState:
    """Device-side tables of the six injection sites + the generator state; the launch plan calls ``coef`` / ``inject``."""

    def __init__(self, model: "CrossFormerWithNoiseB200", device):
        self.encoder = model.encoder_noise
        self.correlated = model.correlated
        self.D = model.noise_latent_dim
        self.seed = int(model.noise_seed)
        self.step = torch.zeros(1, dtype=torch.int64, device=device)   # uint64 counter, advanced once per forward
        self.sites = {}
        sd = {k: v.detach() for k, v in model.state_dict().items()}
        for i, name in enumerate(SITES):
            if name + ".modulation" not in sd:
                continue
            W = sd[name + ".noise_transform.weight"].float().to(device).contiguous()
            self.sites[i] = dict(W=W, b=sd[name + ".noise_transform.bias"].float().to(device).contiguous(),
                                 mod=sd[name + ".modulation"].float().reshape(-1).to(device).contiguous(),
                                 f=sd[name + ".noise_factor"].float().reshape(-1).to(device).contiguous(), C=W.shape[0], coef={})
        self.recorded = None  # {"latent": {site: [B, D]}, "eps": {site: pixel-major [B*HW, C]}}: the reference's draws (tests)

    def coef(self, site: int, B: int):
        s = self.sites[site]
        buf = s["coef"].get(B)
        if buf is None:
            buf = s["coef"][B] = torch.empty((B, s["C"]), device=s["W"].device, dtype=torch.float32)
        lat = None if self.recorded is None else self.recorded["latent"][site]
        # correlated: every site reads the latent stream of site 0 (one draw per forward)
        ops.noise_coef(lat, s["W"], s["b"], s["mod"], s["f"], buf, B, s["C"], self.D, self.seed, self.step,
                       0 if self.correlated else site)

    def inject(self, site, x, ldx, out, ldo, hi, lo, ldh, B, HW, C):
        eps = None if self.recorded is None else self.recorded["eps"][site]
        ops.noise_inject(x, ldx, out, ldo, hi, lo, ldh, 0, self.sites[site]["coef"][B], eps, B, HW, C, self.seed, self.step, site)

    def advance(self):
        ops.noise_step_advance(self.step)


class CrossFormerWithNoiseB200(CrossFormerB200):
    """Constructor = reference keywords (crossformer_ensemble.py:21-30) + ``noise_seed`` (member seed of the generator)."""

    def __init__(self, noise_latent_dim=128, encoder_noise_factor=0.05, decoder_noise_factor=0.275, encoder_noise=True,
                 freeze=True, correlated=False, noise_seed: int = 0, **kwargs):
        super().__init__(**kwargs)
        geo = self.geometry
        if geo.frames != 1:
            raise NotImplementedError("the ensemble variant averages input frames (avg_pool3d, crossformer_ensemble.py:131-132); "
                                      "only frames == 1 is built")
        if geo.variant != "crossformer":
            raise NotImplementedError("CrossFormerWithNoise derives from credit.models.crossformer.CrossFormer")
        self.noise_latent_dim, self.encoder_noise, self.correlated = int(noise_latent_dim), bool(encoder_noise), bool(correlated)
        self.noise_seed = int(noise_seed)
        dims = geo.dim
        chans = {"encoder_noise_layers.0": dims[0], "encoder_noise_layers.1": dims[1], "encoder_noise_layers.2": dims[2],
                 "noise_inject1": geo.ups[0].c_out, "noise_inject2": geo.ups[1].c_out, "noise_inject3": geo.ups[2].c_out}
        g = torch.Generator().manual_seed(self._init_seed + 1)
        for name in SITES:
            if name.startswith("encoder") and not self.encoder_noise:
                continue
            C = chans[name]
            mod = self
            for p in name.split("."):
                if p not in mod._modules:
                    mod.add_module(p, _Holder())
                mod = mod._modules[p]
            factor = encoder_noise_factor if name.startswith("encoder") else decoder_noise_factor
            bound = 1.0 / (self.noise_latent_dim ** 0.5)  # nn.Linear default init
            mod.register_parameter("modulation", nn.Parameter(torch.ones(1, C, 1, 1), requires_grad=False))
            mod.register_parameter("noise_factor", nn.Parameter(torch.tensor([float(factor)]), requires_grad=False))
            lin = _Holder()
            mod.add_module("noise_transform", lin)
            lin.register_parameter("weight", nn.Parameter((torch.rand(C, self.noise_latent_dim, generator=g) * 2 - 1) * bound,
                                                          requires_grad=False))
            lin.register_parameter("bias", nn.Parameter((torch.rand(C, generator=g) * 2 - 1) * bound, requires_grad=False))
        self._noise_state: Optional[NoiseState] = None
        self._recorded = None

    # the base class materialises / loads only the CrossFormer keys lazily; the noise parameters above are eager
    def _materialise(self):
        if not self._lazy_init:
            return
        from .synth import synthetic_state_dict

        self._lazy_init = False
        init = synthetic_state_dict(self.geometry, seed=self._init_seed, sn_iters=5)
        with torch.no_grad():
            own = nn.Module.state_dict(self)
            for k, v in init.items():
                own[k].copy_(v)

    def set_recorded_noise(self, draws: Optional[List[torch.Tensor]]):
        """Feed the reference's recorded ``torch.randn`` draws (tests/golden/make_golden_ensemble.py) instead of the generator:
        NCHW eps tensors are re-laid pixel-major.  ``None`` switches back to the Philox generator."""
        self._recorded = draws
        self._noise_state = None
        self._plans = {}

    def _recorded_tables(self, device):
        if self._recorded is None:
            return None
        draws = list(self._recorded)
        sites = [i for i, n in enumerate(SITES) if (self.encoder_noise or not n.startswith("encoder"))]
        lat, eps = {}, {}
        if self.correlated:
            one = draws.pop(0).to(device).contiguous()
            for i in sites:
                lat[i] = one
                e = draws.pop(0)
                eps[i] = e.permute(0, 2, 3, 1).reshape(-1, e.shape[1]).to(device).contiguous()
        else:
            for i in sites:
                lat[i] = draws.pop(0).to(device).contiguous()
                e = draws.pop(0)
                eps[i] = e.permute(0, 2, 3, 1).reshape(-1, e.shape[1]).to(device).contiguous()
        return {"latent": lat, "eps": eps}

    def refresh_weights(self):
        super().refresh_weights()
        self._noise_state = None

    def _plan_for(self, x: torch.Tensor):
        if self._domain is not None:
            raise NotImplementedError("the ensemble variant runs on the single-GPU plan (members are independent forecasts)")
        if self.training:
            raise NotImplementedError("eval-mode forecast forward only: call .eval()")
        geo = self.geometry
        if x.dim() != 5 or tuple(x.shape[1:]) != geo.in_shape:
            raise ValueError(f"expected input [B, {', '.join(map(str, geo.in_shape))}], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("CrossFormerWithNoiseB200 has no CPU path: pass a CUDA tensor")
        if self._prepared is None or self._prepared_sig != self._signature():
            self.refresh_weights()
        if self._noise_state is None:
            self._noise_state = NoiseState(self, x.device)
            self._noise_state.recorded = self._recorded_tables(x.device)
            self._plans = {}
        key = (int(x.shape[0]), x.device.index)
        plan = self._plans.get(key)
        if plan is None:
            with torch.cuda.device(x.device):
                plan = self._plans[key] = _Plan(geo, self._prepared, int(x.shape[0]), x.device, True, noise=self._noise_state)
        return x.float().contiguous(), plan

    @torch.no_grad()
    def forward(self, x: torch.Tensor, noise=None, forecast_step=None) -> torch.Tensor:
        # `noise` is accepted and ignored like in the reference, which overwrites it before every use (:121-122, 142-143)
        x, plan = self._plan_for(x)
        with torch.cuda.device(x.device):
            return plan.run(x)
