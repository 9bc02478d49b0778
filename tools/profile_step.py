"""One forecast step of the headline workload between cudaProfilerStart/Stop (for ncu --profile-from-start off).

Also writes gpurun_out/step_tags.json: the plan's launch tags in order with the number of kernels each enqueues, so that
tools/ncu_traffic.py can attribute the ncu launch list (kernel names only) to the bench's kernel families."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miles_credit_b200 import ops  # noqa: E402
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
plan = next(iter(model._plans.values()))
# launch tags of one forward, in order, with the kernels each entry point enqueues (counted on an unprofiled pass)
tags = []
n0 = ops.LAUNCHES
plan._pad(x)
tags.append(["pad", ops.LAUNCHES - n0])
for fn, args, tag, _fl, _by in plan.steps:
    n0 = ops.LAUNCHES
    fn(*args)
    tags.append([tag, ops.LAUNCHES - n0])
out = torch.empty((1, geo.base_output_channels, geo.output_frames, geo.h_out, geo.w_out), device="cuda")
n0 = ops.LAUNCHES
plan._unpad(out)
tags.append(["unpad_resize", ops.LAUNCHES - n0])
torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(tags, open(os.path.join(ROOT, "gpurun_out", "step_tags.json"), "w"))
torch.cuda.profiler.start()
y = model(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches per forward:", sum(n for _, n in tags), "finite:", bool(torch.isfinite(y).all()))
