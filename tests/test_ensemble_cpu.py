"""CPU: the noise-injection ensemble variant (miles_credit_b200/ensemble.py) — state-dict layout and launch plan through
the C-ABI emulator, fed the recorded draws of the UNMODIFIED reference CrossFormerWithNoise
(tests/golden/make_golden_ensemble.py)."""
import os

import pytest
import torch

from miles_credit_b200 import lib as wlib
from miles_credit_b200 import model as wmodel
from miles_credit_b200 import ops
from miles_credit_b200.ensemble import CrossFormerWithNoiseB200, NoiseState
from miles_credit_b200.geometry import build_geometry
from miles_credit_b200.synth import synthetic_state_dict
from miles_credit_b200.weights import prepare

from abi_emulator import EmulatedLib


@pytest.fixture
def emulated(monkeypatch):
    emu = EmulatedLib()
    monkeypatch.setattr(wlib, "_lib", emu)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(ops, "_req", lambda *a, **k: None)
    return emu


@pytest.mark.parametrize("case", ["unit_ensemble", "unit_ensemble_correlated"])
def test_ensemble_plan_through_emulated_abi_matches_reference(golden_dir, emulated, case):
    fx = torch.load(os.path.join(golden_dir, f"{case}.pt"), weights_only=False)
    model = CrossFormerWithNoiseB200(**fx["kwargs"])
    assert {k: list(v.shape) for k, v in model.state_dict().items()} == fx["keys"]      # the reference's key layout
    base_kw = {k: v for k, v in fx["kwargs"].items() if k not in ("noise_latent_dim", "encoder_noise_factor",
                                                                   "decoder_noise_factor", "encoder_noise", "freeze", "correlated")}
    geo = build_geometry(**base_kw)
    sd = dict(synthetic_state_dict(geo, seed=fx["seed"]), **fx["noise_state"])
    msg = model.load_state_dict(sd, strict=True)
    assert not msg.missing_keys and not msg.unexpected_keys
    model.set_recorded_noise(fx["draws"])
    ns = NoiseState(model, torch.device("cpu"))
    ns.recorded = model._recorded_tables(torch.device("cpu"))
    wts = prepare({k: v for k, v in model.state_dict().items()}, geo, wmodel._round_up(geo.input_channels, 4))
    plan = wmodel._Plan(geo, wts, fx["x"].shape[0], torch.device("cpu"), True, noise=ns)
    y = plan.run(fx["x"])
    err = float((y - fx["y"]).abs().max() / fx["y"].abs().max())
    print(case, "rel-max vs the reference", err)
    assert err < 2e-5, err
    n_sites = 6 if fx["kwargs"]["encoder_noise"] else 3
    assert emulated.calls.count("noise_inject") == n_sites and emulated.calls.count("noise_coef") == n_sites
    # without noise the same weights give a different answer (the injection is really on the path)
    plain = wmodel._Plan(geo, wts, fx["x"].shape[0], torch.device("cpu"), True).run(fx["x"])
    assert float((plain - fx["y"]).abs().max() / fx["y"].abs().max()) > 1e-3
