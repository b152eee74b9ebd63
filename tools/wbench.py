"""Pure-write and pure-read HBM bandwidth (torch fill_ / sum over 8 GiB), for kernels whose traffic is one-sided
(the interpolator writes 8x what it reads)."""
import torch
x = torch.empty(1 << 31, dtype=torch.int32, device="cuda")  # 8 GiB
for name, fn in (("write (fill_)", lambda: x.fill_(7)), ("read (sum)", lambda: x.sum())):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {x.numel() * 4 * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9:.0f} GB/s")
