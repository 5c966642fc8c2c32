"""HBM write-only / read-only / copy rates on this GPU with torch (fill_, sum, copy_) over 1 GiB: the ceilings the
store-heavy decoder phase and the step's output traffic are compared with."""
import torch
n = 1 << 28
a = torch.empty(n, device="cuda", dtype=torch.float32)
b = torch.empty(n, device="cuda", dtype=torch.float32)
def t(f, bytes_, name):
    for _ in range(3): f()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name}: {bytes_ / best / 1e6:.0f} GB/s ({best:.3f} ms)")
t(lambda: a.fill_(1.0), 4 * n, "fill (write only)")
t(lambda: a.sum(), 4 * n, "sum (read only)")
t(lambda: b.copy_(a), 8 * n, "copy (read + write)")
