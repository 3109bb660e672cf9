"""pairwise_iou timing at several sizes / overlap densities (CUDA events, output buffer reused through the allocator)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))
import torch
import det_b200 as det
def boxes(m, frame, size, seed):
    g = torch.Generator().manual_seed(seed)
    xy = torch.rand(m, 2, generator=g) * 0.8 * frame
    wh = torch.rand(m, 2, generator=g) * size + 1
    return torch.cat([xy, xy + wh], 1).cuda()
for m, frame, size in ((20000, 1024.0, 0.2 * 1024), (10000, 1024.0, 0.2 * 1024), (20000, 1024.0, 600.0), (20000, 8192.0, 100.0)):
    b1, b2 = det.Boxes(boxes(m, frame, size, 4)), det.Boxes(boxes(m, frame, size, 5))
    q = det.pairwise_iou(b1, b2)
    dens = float((q > 0).float().mean())
    del q
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        det.pairwise_iou(b1, b2)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    by = 32 * m + 4 * m * m
    print(f"m={m} frame={frame} size<={size:.0f}: overlap density {dens:.3f}  {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s  frac {by/ms/1e6/6537.6:.3f}")
