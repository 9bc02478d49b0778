"""Measured error table for the f16x2 precision scheme: what happens to the forward's rel-max error
when one of the three fp16 passes  A_hi W_hi + A_hi W_lo + A_lo W_hi  is dropped in a family of GEMM launches.

Runs the real launch plan of a WXFormer-6h architecture on the CPU through the C-ABI emulator (tests/abi_emulator.py: the
documented semantics of every entry point on the real operand planes) and compares with the fp32 oracle.  "alo" = the
A_lo W_hi pass dropped (activations effectively rounded to fp16), "wlo" = the A_hi W_lo pass dropped (weights rounded).

    python tools/precision_table.py [workload] > profiles/r2_precision_table.txt
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from abi_emulator import EmulatedLib  # noqa: E402
from miles_credit_b200 import lib as wlib  # noqa: E402
from miles_credit_b200 import model as wmodel  # noqa: E402
from miles_credit_b200 import ops  # noqa: E402
from miles_credit_b200.geometry import build_geometry, workload  # noqa: E402
from miles_credit_b200.synth import synthetic_input, synthetic_state_dict  # noqa: E402
from miles_credit_b200.weights import prepare  # noqa: E402
from oracle import crossformer_oracle as oracle  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "wxformer_6h_1deg"
kw = workload(name)
emu = EmulatedLib()
wlib._lib = emu
ops._stream = lambda: 0
ops._req = lambda *a, **k: None
geo = build_geometry(**kw)
sd = synthetic_state_dict(geo, seed=1000, sn_iters=5)
wts = prepare(sd, geo, wmodel._round_up(geo.input_channels, 4))
plan = wmodel._Plan(geo, wts, 1, torch.device("cpu"), True)
x = synthetic_input(geo, batch=1, seed=1000)
with torch.no_grad():
    ref = oracle.forward(x, sd, geo)
    ref64 = oracle.forward(x.double(), {k: v.double() for k, v in sd.items()}, geo) if os.environ.get("WXF_REF64") else None


def run(rule):
    """rule(tag) -> set of passes to drop for that launch."""
    plan._pad(x)
    for fn, args, tag, _fl, _by in plan.steps:
        emu.drop = rule(tag)
        fn(*args)
    emu.drop = set()
    out = torch.empty((1, *geo.out_shape))
    plan._unpad(out)
    return float((out - ref).abs().max() / ref.abs().max())


fams = ["qkv", "out_proj", "ff1", "ff2"]
print(f"workload {name}: {sum(geo.depth)} transformer blocks, rel-max error of the forward vs the fp32 oracle (tolerance 1e-4)")
print(f"{'three passes everywhere (the shipped scheme)':58s} {run(lambda t: set()):.3e}")
for what in ("alo", "wlo"):
    label = "A_lo W_hi dropped" if what == "alo" else "A_hi W_lo dropped"
    print(f"{label + ' in every GEMM launch':58s} {run(lambda t, w=what: {w} if t.split('.')[0] in fams else set()):.3e}")
    for f in fams:
        print(f"{label + ' in ' + f + ' (all stages)':58s} {run(lambda t, w=what, f=f: {w} if t.split('.')[0] == f else set()):.3e}")
    for s in range(4):
        print(f"{label + f' in ff2 of stage {s} only':58s} {run(lambda t, w=what, s=s: {w} if t == f'ff2.s{s}' else set()):.3e}")
