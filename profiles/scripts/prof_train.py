"""ncu driver: reference-native RPN training step (BASELINE configs[2]-ii): R=50127 anchors, batch N (default 1024):
match (2 kernels) + subsample + fused loss fwd/bwd."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))
import torch
import det_b200 as det
from bench import synth_gt
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
rpn = det.RegionProposalNetwork([4, 8, 16, 32, 64])
anchors = torch.cat(rpn.anchor_generator.grid_anchors([(448 // s, 448 // s) for s in (4, 8, 16, 32, 64)], dev), 0)
R = anchors.shape[0]
gtb, _, off, tot = synth_gt(nb, 3, dev)
logits = torch.randn(nb, R, device=dev); deltas = torch.randn(nb, R, 4, device=dev) * 0.5
gl, gd = torch.empty_like(logits), torch.empty_like(deltas)
rpn._anchors_for_loss = anchors
for it in range(3):
    matched, labels = rpn.anchor_matcher.match_packed(gtb, off, nb, anchors)
    det.subsample_labels_(labels, 256, 0.5, it)
    asg = det.Assignment(labels, matched, gtb, off)
    rpn._run_loss(logits, deltas, asg, nb, None, gl, gd)
torch.cuda.synchronize()
print("ok", R, tot)
