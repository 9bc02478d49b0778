"""One forecast step of the headline workload between cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miles_credit_b200.geometry import build_geometry, workload  # noqa: E402
from miles_credit_b200.model import CrossFormerB200  # noqa: E402
from miles_credit_b200.synth import synthetic_input, synthetic_state_dict  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "wxformer_6h_025deg"
kw = workload(name)
geo = build_geometry(**kw)
model = CrossFormerB200(**kw)
model.load_state_dict(synthetic_state_dict(geo, seed=1000, sn_iters=3), strict=True)
model = model.cuda().eval()
x = synthetic_input(geo, batch=1, seed=1000).cuda()
for _ in range(2):
    model(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
y = model(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches per forward:", len(next(iter(model._plans.values())).steps) + 2, "finite:", bool(torch.isfinite(y).all()))
