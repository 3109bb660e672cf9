"""ncu driver: dense anchor head decode (BASELINE configs[3] geometry), batch 32, inputs rotate over > L2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))
import torch
import det_b200 as det
n, C = (int(sys.argv[1]) if len(sys.argv) > 1 else 32), 80
strides = [8, 16, 32]
wh = [[[10, 13], [16, 30], [33, 23]], [[30, 61], [62, 45], [59, 119]], [[116, 90], [156, 198], [373, 326]]]
dh = det.DenseAnchorHead(strides, wh, C)
heads = [[torch.randn(n, 3 * (5 + C), 640 // s, 640 // s, device="cuda") for s in strides] for _ in range(3)]
out = None
for i in range(6):
    out = dh.decode(heads[i % 3], out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(30):
    dh.decode(heads[i % 3], out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 30
print(f"N={n}: ms/decode {ms:.4f}  {n * 9273600 / ms / 1e6:.0f} GB/s algorithmic")
