"""A/B of programmatic dependent launch on back-to-back launches of the fused grid-head kernel (run twice: DET_NO_PDL=0/1).
Times (a) eager launches on one stream, (b) one CUDA graph of 95 launches, (c) the ncu-visible single launch is separate."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))
import torch
import det_b200 as det
yh = det.YoloGridHead(7, 2, 20, (448, 448))
pool = 95
heads = torch.randn(pool, 256, 7, 7, 30, generator=torch.Generator().manual_seed(1)).cuda()
outs = [yh.detect(heads[i], 0.25, 0.5, max_det=300) for i in range(4)]
torch.cuda.synchronize()
def run():
    for i in range(pool):
        yh.detect(heads[i], 0.25, 0.5, max_det=300, out=outs[i % 4])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): run()
torch.cuda.synchronize(); e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
eager = e0.elapsed_time(e1) / (20 * pool) * 1e3
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side): run()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): run()
for _ in range(3): g.replay()
torch.cuda.synchronize(); e0.record()
for _ in range(50): g.replay()
e1.record(); torch.cuda.synchronize()
graph = e0.elapsed_time(e1) / (50 * pool) * 1e3
print(f"DET_NO_PDL={os.environ.get('DET_NO_PDL', '0')}: eager {eager:.2f} us/launch, graph {graph:.2f} us/launch")
