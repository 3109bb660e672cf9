"""Where does the end-to-end step go?  Raw pinned H2D / D2H bandwidth at the step's sizes, then YoloHostPipeline at
several depths, with and without the per-step host read of the result."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))
import torch
import det_b200 as det

dev = torch.device("cuda", 0)
def bw(nbytes, h2d, iters=200):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(5):
        (d.copy_(h, non_blocking=True) if h2d else h.copy_(d, non_blocking=True))
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        (d.copy_(h, non_blocking=True) if h2d else h.copy_(d, non_blocking=True))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms * 1e3, nbytes / ms / 1e6
for nb, h2d in ((1505280, True), (2151424, False), (64 << 20, True), (64 << 20, False)):
    us, gbs = bw(nb, h2d)
    print(f"{'H2D' if h2d else 'D2H'} {nb:>9d} B: {us:8.1f} us  {gbs:6.1f} GB/s")
yh = det.YoloGridHead(7, 2, 20, (448, 448))
for depth in (2, 3, 4, 8):
    for read in (True, False):
        pipe = det.YoloHostPipeline(yh, 256, 0.25, 0.5, 300, depth=depth)
        src = torch.randn(depth, 256, 7, 7, 30)
        for s in range(depth):
            pipe.input(s).copy_(src[s])
        def run(k):
            for i in range(k):
                slot = i % depth
                if i >= depth:
                    r = pipe.wait(slot)
                    if read:
                        int(r["count"][0])
                pipe.launch(slot)
            for s in range(depth):
                pipe.wait(s)
        run(50); torch.cuda.synchronize()
        t0 = time.perf_counter(); run(2000); dt = time.perf_counter() - t0
        print(f"depth {depth} read={read}: {dt / 2000 * 1e6:7.1f} us/step  {256 * 2000 / dt / 1e6:6.2f} M img/s")
# host-side cost of one launch() when nothing has to be waited for
pipe = det.YoloHostPipeline(yh, 256, 0.25, 0.5, 300, depth=3)
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(300):
    pipe.launch(i % 3)
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"launch() host cost: {(t1 - t0) / 300 * 1e6:.1f} us")
