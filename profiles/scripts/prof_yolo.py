"""Tiny driver for ncu: a few launches of the fused YOLO decode+NMS kernel on the BASELINE configs[1] shape."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))
import torch
import det_b200 as det

yh = det.YoloGridHead(7, 2, 20, (448, 448))
heads = torch.randn(6, 256, 7, 7, 30, generator=torch.Generator().manual_seed(1)).cuda()
out = None
for i in range(6):
    out = yh.detect(heads[i], 0.25, 0.5, max_det=300, out=out)
torch.cuda.synchronize()
print("kept/img", float(out["count"].float().mean()))
