"""Read-only HBM bandwidth reference points on this GPU (development aid): torch reductions over 1 GiB."""
import torch

x = torch.empty(1 << 28, dtype=torch.int32, device="cuda").random_(0, 100)   # 1 GiB
y = torch.empty_like(x)
def t(fn, nbytes, label):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{label}: {best:.3f} ms  {nbytes / best / 1e6:.0f} GB/s")
t(lambda: x.sum(), x.numel() * 4, "int32 sum (read-only, 1 GiB)")
t(lambda: x.max(), x.numel() * 4, "int32 max (read-only, 1 GiB)")
xf = x.view(torch.float32)
t(lambda: xf.sum(), x.numel() * 4, "fp32 sum (read-only, 1 GiB)")
t(lambda: y.copy_(x), x.numel() * 8, "copy (read + write, 2 GiB moved)")
t(lambda: y.zero_(), x.numel() * 4, "memset (write-only, 1 GiB)")
