"""Reference points for the HBM roofline on this box: read-only and copy streams through stock torch kernels."""
import torch
dev = "cuda"
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it
for mb in (297, 1024, 4096):
    n = mb * 1024 * 1024 // 4
    xs = [torch.randn(n, device=dev) for _ in range(3 if mb < 2000 else 2)]
    y = torch.empty_like(xs[0])
    st = {"i": 0}
    def rd():
        st["i"] += 1; return xs[st["i"] % len(xs)].sum()
    def cp():
        st["i"] += 1; y.copy_(xs[st["i"] % len(xs)])
    def mx():
        st["i"] += 1; return xs[st["i"] % len(xs)].view(-1, 1024).amax(dim=1)
    a, b, c = t(rd), t(cp), t(mx)
    print(f"{mb} MB: sum {n*4/a/1e6:7.0f} GB/s | amax {n*4/c/1e6:7.0f} GB/s | copy {2*n*4/b/1e6:7.0f} GB/s (r+w)")
