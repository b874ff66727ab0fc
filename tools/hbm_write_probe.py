"""HBM write-only / read-only / copy bandwidth on this GPU (torch fill_, sum, copy_ on 4 GB)."""
import torch
x = torch.empty(1 << 30, dtype=torch.float32, device="cuda")   # 4 GB
y = torch.empty_like(x)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
gb = x.numel() * 4 / 1e9
ms = t(lambda: x.fill_(1.0)); print(f"write-only (fill_): {gb/ms*1e3:.0f} GB/s")
ms = t(lambda: x.sum()); print(f"read-only (sum): {gb/ms*1e3:.0f} GB/s")
ms = t(lambda: y.copy_(x)); print(f"copy (read+write counted): {2*gb/ms*1e3:.0f} GB/s")
