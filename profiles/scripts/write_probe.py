"""Pure-write bandwidth of the B200 for different payloads: is a write-only kernel's roofline the copy peak?"""
import torch
dev = torch.device("cuda")
n = 400_000_000  # 1.6 GB of fp32, the 20 000 x 20 000 IoU matrix
x = torch.empty(n, dtype=torch.float32, device=dev)
row = torch.rand(4096, device=dev)
rows = torch.rand(1 << 20, device=dev)
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
for name, fn in (("zero_", lambda: x.zero_()), ("fill_(1.5)", lambda: x.fill_(1.5)),
                 ("broadcast 16 KB random row", lambda: x[: (n // 4096) * 4096].view(-1, 4096).copy_(row)),
                 ("broadcast 4 MB random block", lambda: x[: (n // (1 << 20)) * (1 << 20)].view(-1, 1 << 20).copy_(rows)),
                 ("sparse: zeros, 5% random values", None)):
    if fn is None:
        src = torch.where(torch.rand(1 << 20, device=dev) < 0.05, torch.rand(1 << 20, device=dev), torch.zeros(1 << 20, device=dev))
        fn = lambda: x[: (n // (1 << 20)) * (1 << 20)].view(-1, 1 << 20).copy_(src)
    ms = t(fn)
    print(f"{name:34s}: {ms*1e3:7.1f} us  {4*n/ms/1e6:7.0f} GB/s written")
y = torch.empty_like(x)
ms = t(lambda: y.copy_(x))
print(f"{'copy 1.6 GB -> 1.6 GB':34s}: {ms*1e3:7.1f} us  {8*n/ms/1e6:7.0f} GB/s read+written")
ms = t(lambda: torch.arange(n, out=x, dtype=torch.float32))
print(f"{'arange(out=x) (distinct values)':34s}: {ms*1e3:7.1f} us  {4*n/ms/1e6:7.0f} GB/s written")
xi = torch.empty(n, dtype=torch.int32, device=dev)
ms = t(lambda: torch.arange(n, out=xi, dtype=torch.int32))
print(f"{'arange int32 (distinct values)':34s}: {ms*1e3:7.1f} us  {4*n/ms/1e6:7.0f} GB/s written")
h = torch.empty(n // 2, dtype=torch.float32, device=dev)
ms = t(lambda: torch.add(x[: n // 2], 1.0, out=h))
print(f"{'add 0.8 GB -> 0.8 GB':34s}: {ms*1e3:7.1f} us  {4*n/ms/1e6:7.0f} GB/s read+written")
