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
torch.cuda.profiler.start()
y = model(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("finite:", bool(torch.isfinite(y).all()))
