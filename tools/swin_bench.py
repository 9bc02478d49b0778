"""FuXi Swin window attention of one block at 0.25 deg timed alone (CUDA events, L2 flushed): python tools/swin_bench.py [--lib path]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=None)
args = ap.parse_args()
from miles_credit_b200 import lib as wlib  # noqa: E402

if args.lib:
    wlib.load(os.path.join(ROOT, args.lib))
from miles_credit_b200 import ops  # noqa: E402

dev = "cuda"
H, W, d, heads, ws = 105, 203, 1024, 8, (7, 7)
torch.manual_seed(0)
qkv = torch.randn(H * W, 3 * d, device=dev)
bias = torch.randn(heads, 49, 49, device=dev)
scale = torch.rand(heads, device=dev) + 1.0
hi = torch.empty(H * W, d, device=dev, dtype=torch.float16)
lo = torch.empty_like(hi)
flush = torch.empty(64 << 20, device=dev)
for shift in ((0, 0), (3, 3)):
    ts = []
    for _ in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.swin_window_attention(qkv, 3 * d, bias, scale, hi, lo, None, d, 1, H, W, d, heads, ws, shift)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"{args.lib or 'default':22s} shift {shift}: median {ts[len(ts) // 2]:7.1f} us  min {ts[0]:7.1f}  checksum {float(hi.float().sum()):.4f}")
