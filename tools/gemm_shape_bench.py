"""One GEMM shape of the forecast step timed alone (CUDA events, L2 flushed between launches), optionally through a
diagnosis build of the library: python tools/gemm_shape_bench.py [--lib tools/ablate/lib_a1.so] M N K mode ...
mode: planes | gelu | f32 | red"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=None)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("shapes", nargs="+", help="M,N,K,mode")
args = ap.parse_args()
from miles_credit_b200 import lib as wlib  # noqa: E402

if args.lib:
    wlib.load(os.path.join(ROOT, args.lib))
from miles_credit_b200 import ops  # noqa: E402
from miles_credit_b200.weights import gemm_weights  # noqa: E402

dev = "cuda"
flush = torch.empty(64 << 20, device=dev)
for spec in args.shapes:
    m, n, k, mode = spec.split(",")
    m, n, k = int(m), int(n), int(k)
    torch.manual_seed(0)
    a = torch.randn(m, k, device=dev)
    a_hi = torch.empty(m, k, device=dev, dtype=torch.float16)
    a_lo = torch.empty_like(a_hi)
    ops.split_f16x2(a, k, a_hi, a_lo, k, m, k)
    gw = gemm_weights(torch.randn(n, k, device=dev) / k**0.5, torch.randn(n, device=dev) * 0.1)
    kw = dict(M=m, lda=k)
    if mode in ("planes", "gelu"):
        o_hi = torch.empty(m, n, device=dev, dtype=torch.float16)
        o_lo = torch.empty_like(o_hi)
        kw.update(out_hi=o_hi, out_lo=o_lo, ldh=n, act=(wlib.ACT_GELU if mode == "gelu" else wlib.ACT_NONE))
    else:
        out = torch.zeros(m, n, device=dev)
        kw.update(out=out, ldc=n)
        if mode == "red":
            kw.update(res=out, ldr=n)
    desc = ops.make_gemm_desc(a_hi, a_lo, gw, **kw)
    ops.gemm_f16x2_tc(desc)
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm_f16x2_tc(desc)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"{args.lib or 'default':24s} {spec:24s} median {med:8.1f} us  min {ts[0]:8.1f}  {2.0 * m * n * k / med / 1e6:7.1f} TF/s algorithmic")
