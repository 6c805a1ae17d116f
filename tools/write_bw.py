"""Pure-write and copy bandwidth of this GPU (torch fill / copy on 1 GiB), for the write-bound kernels' roofline."""
import torch

n = 1 << 28  # 1 GiB of float32
x = torch.empty(n, device="cuda")
y = torch.empty(n, device="cuda")


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


t = timed(lambda: x.fill_(1.0))
print(f"fill  1 GiB: {t:.3f} ms  {4 * n / t / 1e6:.0f} GB/s written")
t = timed(lambda: y.copy_(x))
print(f"copy  1 GiB: {t:.3f} ms  {8 * n / t / 1e6:.0f} GB/s read+written ({4 * n / t / 1e6:.0f} GB/s each way)")
t = timed(lambda: x.sum())
print(f"read  1 GiB: {t:.3f} ms  {4 * n / t / 1e6:.0f} GB/s read")
