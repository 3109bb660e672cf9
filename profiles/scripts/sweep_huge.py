import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/object-detection-pytorch-rust_b200")
import torch, torchvision
import det_b200 as det
from oracle import ref_torch as O
dev = torch.device("cuda", 0)
def t(fn, it):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it
NIMG = 8
for M in (5000, 10000, 20000, 50000, 100000):
    g = torch.Generator().manual_seed(4)
    xy = torch.rand(NIMG, M, 2, generator=g) * 0.8 * 1024; wh = torch.rand(NIMG, M, 2, generator=g) * 0.2 * 1024 + 1
    boxes = torch.cat([xy, xy + wh], 2)
    scores = torch.stack([(torch.randperm(M, generator=g).float() + 0.5) / M for _ in range(NIMG)])
    for ncat, maxout in ((1, M), (1, 1000), (3, M), (80, M)):
        cats = torch.randint(0, ncat, (NIMG, M), generator=g)
        b, s, c = boxes.to(dev), scores.to(dev), cats.to(dev)
        keep, cnt = det.nms_images(b, s, c, None, 0.5, maxout, mode=1)
        torch.cuda.synchronize()
        ms = t(lambda: det.nms_images(b, s, c, None, 0.5, maxout, mode=1), 3)
        ref = [torchvision.ops.batched_nms(b[i], s[i], c[i], 0.5)[:maxout] for i in range(NIMG)]
        ms_tv = t(lambda: [torchvision.ops.batched_nms(b[i], s[i], c[i], 0.5) for i in range(NIMG)], 2)
        same = all(int(cnt[i]) == ref[i].numel() and torch.equal(keep[i, :ref[i].numel()], ref[i]) for i in range(NIMG))
        print(f"M={M} cats={ncat} max_out={maxout}: ours {ms:9.3f} ms  torchvision-cuda {ms_tv:9.3f} ms  x{ms_tv/ms:6.1f}  kept {float(cnt.float().mean()):.0f} equal={same}", flush=True)
