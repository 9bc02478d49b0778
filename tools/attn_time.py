"""Times the window-attention launches of the headline workload's four stages in isolation (CUDA events)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:  # python tools/attn_time.py [path of an alternative libwxformer_b200.so]
    from miles_credit_b200 import lib as wlib
    wlib.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), sys.argv[1]))
from miles_credit_b200 import ops

torch.manual_seed(0)
cases = [(0, 400, 800, 128, 10, 0), (0, 400, 800, 128, 10, 1), (1, 200, 400, 256, 10, 0), (1, 200, 400, 256, 5, 1),
         (2, 100, 200, 512, 10, 0), (2, 100, 200, 512, 2, 1), (3, 50, 100, 1024, 10, 0)]
for s, h, w, d, wsz, kind in cases:
    m = h * w
    q_hi = (torch.randn(m, 3 * d, device="cuda") * 0.5).half()
    q_lo = (torch.randn(m, 3 * d, device="cuda") * 1e-4).half()
    o_hi = torch.zeros(m, d, device="cuda", dtype=torch.float16)
    o_lo = torch.zeros_like(o_hi)
    L = wsz * wsz
    tile = ops.attention_bias_tile((torch.randn(L, L, device="cuda") * 0.5).contiguous(), w, wsz, kind)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.window_attention_tc(q_hi, q_lo, 3 * d, tile, o_hi, o_lo, d, 1, h, w, d, 32, wsz, kind, 32 ** -0.5)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"s{s} kind={kind} wsz={wsz}: {min(ts[1:]):.1f} us (median {sorted(ts[1:])[2]:.1f})", flush=True)
