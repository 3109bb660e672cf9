"""ncu driver: per-class NMS over the 25 200 decoded boxes of a dense head, 32 images (BASELINE configs[3])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))
import torch
import det_b200 as det
n, R = 32, 25200
g = torch.Generator().manual_seed(3)
xy = torch.rand(n, R, 2, generator=g) * 600; wh = torch.rand(n, R, 2, generator=g) * 120 + 2
boxes = torch.cat([xy, xy + wh], 2).cuda()
scores = torch.rand(n, R, generator=g).cuda(); classes = torch.randint(0, 80, (n, R), generator=g).cuda()
for _ in range(3):
    keep, cnt = det.nms_images(boxes, scores, classes, None, 0.5, 1000)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): det.nms_images(boxes, scores, classes, None, 0.5, 1000)
e1.record(); torch.cuda.synchronize()
print("ms", e0.elapsed_time(e1) / 10, "kept", float(cnt.float().mean()))
