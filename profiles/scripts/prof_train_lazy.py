"""ncu driver: the training step RegionProposalNetwork.forward(training) launches (BASELINE configs[2]-ii, R=50127, batch N,
default 1024): det_assign_sampled (row maxima -> gt-centric positive search -> lazy sampler) -> det_rpn_loss_sampled."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))
import torch
import det_b200 as det
from bench import synth_gt
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
strides = (4, 8, 16, 32, 64)
rpn = det.RegionProposalNetwork(list(strides))
hw = [(448 // s, 448 // s) for s in strides]
anchors = torch.cat(rpn.anchor_generator.grid_anchors(hw, dev), 0)
grid = rpn.anchor_generator.grid_layout(hw)
gtb, _, off, tot = synth_gt(nb, 3, dev)
obj = [torch.randn(nb, 3, h, w, device=dev) for h, w in hw]
dlt = [torch.randn(nb, 12, h, w, device=dev) * 0.5 for h, w in hw]
g_obj, g_dlt = [torch.zeros_like(o) for o in obj], [torch.zeros_like(d) for d in dlt]
prev = None
for it in range(iters):
    asg = rpn.assign_sampled(anchors, gtb, off, nb, grid, seed=it + 1)
    rpn._run_sampled(anchors, obj, dlt, asg, nb, None, (g_obj, g_dlt), prev)
    prev = (asg.samples, asg.sample_count)
torch.cuda.synchronize()
print("ok", anchors.shape[0], tot, float(asg.sample_count.float().mean()))
