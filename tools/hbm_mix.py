"""HBM bandwidth by read/write mix (development aid): torch fill_ (write only), copy_ (1:1),
a 2-read-1-write add and a 1-read-2-write split, to place the attention kernels' 39/61 (forward)
and 67/33 (backward) read/write mixes against the copy figure in MEASURED_PEAKS.json."""
import torch
n = 1 << 29                       # 1 GiB of bf16 per tensor
a = torch.empty(n, dtype=torch.bfloat16, device="cuda").normal_()
b = torch.empty_like(a).normal_()
c = torch.empty_like(a)
d = torch.empty_like(a)

def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best
gb = n * 2 / 1e9
print(f"fill_  (0R:1W): {gb / t(lambda: c.fill_(1.0)):.0f} GB/s")
print(f"copy_  (1R:1W): {2 * gb / t(lambda: c.copy_(a)):.0f} GB/s")
print(f"add    (2R:1W): {3 * gb / t(lambda: torch.add(a, b, out=c)):.0f} GB/s")
print(f"sum    (1R:0W): {gb / t(lambda: a.sum()):.0f} GB/s")
