"""ncu driver: pairwise_iou at M x M (default 20000), matrix materialised (BASELINE configs[4])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))
import torch
import det_b200 as det
m = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
g = torch.Generator().manual_seed(4)
xy = torch.rand(2 * m, 2, generator=g) * 0.8 * 1024
wh = torch.rand(2 * m, 2, generator=g) * 0.2 * 1024 + 1
bx = torch.cat([xy, xy + wh], 1).cuda()
b1, b2 = det.Boxes(bx[:m]), det.Boxes(bx[m:])
for _ in range(3):
    q = det.pairwise_iou(b1, b2)
torch.cuda.synchronize()
print("ok", float(q.max()))
