"""ncu driver: single-category NMS of 8 x M boxes (the huge-segment cooperative kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))
import torch
import det_b200 as det
M = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
g = torch.Generator().manual_seed(4)
xy = torch.rand(8, M, 2, generator=g) * 0.8 * 1024; wh = torch.rand(8, M, 2, generator=g) * 0.2 * 1024 + 1
b = torch.cat([xy, xy + wh], 2).cuda()
s = torch.stack([(torch.randperm(M, generator=g).float() + 0.5) / M for _ in range(8)]).cuda()
for _ in range(2):
    keep, cnt = det.nms_images(b, s, None, None, 0.5, M, mode=1)
torch.cuda.synchronize()
print(cnt.tolist())
