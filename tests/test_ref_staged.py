"""CPU: the staged copy of the UNMODIFIED reference (oracle/_ref, recipe oracle/make_ref.py) is the reference.

It must reproduce the committed golden outputs bit for bit (they were generated from /root/reference by
tests/golden/make_golden.py) and its files must hash to the manifest the recipe wrote."""
import hashlib
import json
import os
import subprocess
import sys

import pytest
import torch

from miles_credit_b200.geometry import build_geometry
from miles_credit_b200.synth import synthetic_input, synthetic_state_dict
from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref not staged (run oracle/make_ref.py)")


@needs_ref
def test_manifest_hashes():
    man = json.load(open(os.path.join(ref_loader.STAGED, "MANIFEST.json")))
    assert man["commit"].startswith("5e14b83c")
    assert "credit/models/crossformer.py" in man["files"] and "credit/boundary_padding.py" in man["files"]
    for rel, sha in man["files"].items():
        with open(os.path.join(ref_loader.STAGED, rel), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == sha, rel


@needs_ref
@pytest.mark.parametrize("case", ["unit", "unit_wxformer"])
def test_staged_reference_reproduces_golden_bit_exact(golden_dir, case):
    # a fresh interpreter: nothing but oracle/_ref may provide `credit`
    code = f"""
import sys, torch
sys.path.insert(0, {ROOT!r})
assert not any(p.rstrip('/') == '/root/reference' for p in sys.path)
from oracle import ref_loader
from miles_credit_b200.geometry import build_geometry
from miles_credit_b200.synth import synthetic_input, synthetic_state_dict
fx = torch.load({os.path.join(golden_dir, case + '.pt')!r}, weights_only=False)
geo = build_geometry(**fx['kwargs'])
sd = synthetic_state_dict(geo, seed=fx['seed'])
m = ref_loader.reference_model(fx['kwargs'], sd, fx['kwargs'].get('variant', 'crossformer'))
import credit
assert credit.__file__.startswith({ref_loader.STAGED!r}), credit.__file__
with torch.no_grad():
    y = m(synthetic_input(geo, batch=fx['batch'], seed=fx['seed']))
print('BITEXACT', bool(torch.equal(y, fx['y'])), float((y - fx['y']).abs().max()))
"""
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "BITEXACT True" in res.stdout, res.stdout


@pytest.mark.skipif(not ref_loader.available("fuxi"), reason="oracle/_ref not staged (run oracle/make_ref.py)")
def test_staged_fuxi_reference_reproduces_golden_bit_exact(golden_dir):
    """credit/models/fuxi.py from oracle/_ref (with the Swin-V2 stand-in registered as timm, as when the golden was made)."""
    code = f"""
import sys, torch
sys.path.insert(0, {ROOT!r})
from oracle import ref_loader
fx = torch.load({os.path.join(golden_dir, 'unit_fuxi.pt')!r}, weights_only=False)
m = ref_loader.reference_model(fx['kwargs'], fx['state_dict'], 'fuxi')
import credit.models.fuxi as f
assert f.__file__.startswith({ref_loader.STAGED!r}), f.__file__
with torch.no_grad():
    y = m(fx['x'])
print('BITEXACT', bool(torch.equal(y, fx['y'])), float((y - fx['y']).abs().max()))
"""
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "BITEXACT True" in res.stdout, res.stdout
