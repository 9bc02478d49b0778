"""One FuXi forecast step between cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miles_credit_b200 import fuxi as F  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "fuxi_6h_025deg"
kw = F.fuxi_workload(name)
geo = F.build_fuxi_geometry(**kw)
model = F.FuxiB200(**kw)
model.load_state_dict(F.synthetic_fuxi_state_dict(geo, seed=1000, sn_iters=3), strict=True)
model = model.cuda().eval()
x = F.synthetic_fuxi_input(geo, batch=1, seed=1000).cuda()
for _ in range(2):
    model(x)
torch.cuda.synchronize()
# launch tags of one forward, in order, with the kernels each entry point enqueues (for tools/ncu_traffic.py)
import json  # noqa: E402

from miles_credit_b200 import ops  # noqa: E402

plan = next(iter(model._plans.values()))
tags = []
n0 = ops.LAUNCHES
plan._pad(x)
tags.append(["pad", ops.LAUNCHES - n0])
for fn, args, tag, _fl, _by in plan.steps:
    n0 = ops.LAUNCHES
    fn(*args)
    tags.append([tag, ops.LAUNCHES - n0])
out = torch.empty((1, *geo.out_shape), device="cuda")
n0 = ops.LAUNCHES
plan._unpad(out)
tags.append(["unpad_resize", ops.LAUNCHES - n0])
torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(tags, open(os.path.join(ROOT, "gpurun_out", "step_tags_fuxi.json"), "w"))
torch.cuda.profiler.start()
y = model(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("finite:", bool(torch.isfinite(y).all()))
