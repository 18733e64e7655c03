"""Diagnostic: write-only and copy bandwidth of the device for a 351 MB buffer (the size of one step's feature tensor)."""
import torch
n = 64 * 4 * 2 * 287 * 597
a = torch.empty(n, device="cuda"); b = torch.empty(n, device="cuda")
def timed(fn, k=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
t = timed(lambda: a.fill_(1.0)); print("fill  351 MB: %.4f ms  %.0f GB/s written" % (t, n * 4 / t / 1e6))
t = timed(lambda: a.zero_()); print("zero  351 MB: %.4f ms  %.0f GB/s written" % (t, n * 4 / t / 1e6))
t = timed(lambda: b.copy_(a)); print("copy  351 MB: %.4f ms  %.0f GB/s read+written" % (t, 2 * n * 4 / t / 1e6))
