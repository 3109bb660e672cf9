"""Dense head decode at several batch sizes: the one-launch bulk-copy pipeline vs the per-level LDG kernels vs a stock
torch reduction over the same layout (amax over the class planes), all in algorithmic GB/s."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))
import torch
import det_b200 as det
from det_b200 import _native as N
C = 80
strides = [8, 16, 32]
wh = [[[10, 13], [16, 30], [33, 23]], [[30, 61], [62, 45], [59, 119]], [[116, 90], [156, 198], [373, 326]]]
dh = det.DenseAnchorHead(strides, wh, C)
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it
for n in [int(a) for a in sys.argv[1:]] or [32, 256]:
    sets = 3 if n <= 64 else 2
    heads = [[torch.randn(n, 3 * (5 + C), 640 // s, 640 // s, device="cuda") for s in strides] for _ in range(sets)]
    out = dh.decode(heads[0])
    st = {"i": 0}
    def one():
        st["i"] += 1; dh.decode(heads[st["i"] % sets], out=out)
    def per_level():
        st["i"] += 1; hs = heads[st["i"] % sets]; off = 0
        for h, a, s in zip(hs, dh._anchors_on(h.device if False else hs[0].device), strides):
            N.call("det_dense_decode_level", N.ptr(h), n, 3, C, h.shape[2], h.shape[3], s, N.ptr(a), dh.scale_clamp,
                   N.ptr(out[0]), N.ptr(out[1]), N.ptr(out[2]), 25200, off, N.stream())
            off += h.shape[2] * h.shape[3] * 3
    def stock():
        st["i"] += 1; hs = heads[st["i"] % sets]
        return [h.view(n, 3, 5 + C, h.shape[2], h.shape[3])[:, :, 5:].amax(dim=2) for h in hs]
    by = n * 9273600
    a, b, c = t(one), t(per_level), t(stock)
    print(f"N={n}: one-launch {a*1e3:7.1f} us {by/a/1e6:6.0f} GB/s | per-level {b*1e3:7.1f} us {by/b/1e6:6.0f} GB/s | torch amax {c*1e3:7.1f} us {by/c/1e6:6.0f} GB/s")
