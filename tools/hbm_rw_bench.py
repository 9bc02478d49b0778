"""HBM bandwidth by traffic mix on this GPU (CUDA events, buffers >> L2): pure read, pure write, copy, 1:4 read:write
(the mix of the FeedForward fc1 launch: 164 MB of operand planes in, 655 MB of hidden planes out)."""
import torch

dev = "cuda"
n = 1 << 28  # 1 GiB of fp32
a = torch.empty(n, device=dev)
b = torch.empty(n, device=dev)
a.normal_()


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


gb = n * 4 / 1e9
ms = timed(lambda: b.copy_(a))
print(f"copy        {2 * gb / ms * 1e3:8.0f} GB/s  (read + write)")
ms = timed(lambda: b.fill_(1.0))
print(f"pure write  {gb / ms * 1e3:8.0f} GB/s  (fill)")
ms = timed(lambda: b.zero_())
print(f"memset      {gb / ms * 1e3:8.0f} GB/s")
ms = timed(lambda: a.sum())
print(f"pure read   {gb / ms * 1e3:8.0f} GB/s  (sum)")
q = n // 4
c = torch.empty(n, device=dev)
def mix():
    torch.add(a[:q], 1.0, out=c[:q])
    c[q:].fill_(2.0)


ms = timed(mix)
print(f"1:4 mix     {(q * 4 * 2 + (n - q) * 4) / 1e9 / ms * 1e3:8.0f} GB/s  (two kernels back to back)")
