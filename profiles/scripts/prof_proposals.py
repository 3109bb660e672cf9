"""RPN inference path of the reference (rpn.py:299-328): NCHW head -> decode -> find_top_rpn_proposals, whole batch.
FPN-18 at 448x448 (R = 50127 anchors, 5 levels), pre_nms_topk 2000 / level, post_nms_topk 1000, NMS 0.7.
Prints our time per batch, the CPU oracle's time per image, and checks the proposals of two images against the oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))
import torch
import det_b200 as det
from oracle import ref_torch as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64  # usage: prof_proposals.py [n] [pre] [post] [--clustered]
PRE = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
POST = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
dev = torch.device("cuda", 0)
strides = [4, 8, 16, 32, 64]
rpn = det.RegionProposalNetwork(strides)
g = torch.Generator().manual_seed(6)
obj = [torch.randn(n, 3, 448 // s, 448 // s, generator=g) for s in strides]
dlt = [torch.randn(n, 12, 448 // s, 448 // s, generator=g) * 0.4 for s in strides]
if "--clustered" in sys.argv:
    # objects: logits are a coarse random field (one value per 8x8 block of positions, shared by the cell anchors) plus a
    # little noise, deltas small -- neighbouring anchors score alike and overlap, as the outputs of a trained RPN do, so
    # the NMS suppresses most of the top candidates and the tier cut falls short (full segments are swept)
    for l, s in enumerate(strides):
        hh = 448 // s
        coarse = torch.randn(n, 1, (hh + 7) // 8, (hh + 7) // 8, generator=g)
        field = torch.nn.functional.interpolate(coarse, size=(hh, hh), mode="nearest")
        obj[l] = field.expand(n, 3, hh, hh) * 2.0 + 0.05 * torch.randn(n, 3, hh, hh, generator=g)
        dlt[l] = dlt[l] * 0.1
obj_d, dlt_d = [t.contiguous().to(dev) for t in obj], [t.to(dev) for t in dlt]
sizes = torch.tensor([[448, 448]] * n, dtype=torch.int32, device=dev)
def run():
    logits, boxes, level_sizes = rpn.decode_heads(obj_d, dlt_d)
    return det.rpn_proposals_batched(boxes, logits, level_sizes, sizes, 0.7, PRE, POST, 0.0)
for _ in range(3): out = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): out = run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
from bench import time_graph, load_peaks
ms_dec = time_graph([lambda: rpn.decode_heads(obj_d, dlt_d)], 20)
R = 50127
by = n * 40 * R  # read 4R logits + 16R deltas, write 4R logits + 16R boxes (anchors synthesised: 0 bytes)
print(f"decode_heads (det_rpn_decode, one launch for the five levels): {ms_dec * 1e3:.1f} us = {by / ms_dec / 1e6:.0f} GB/s algorithmic "
      f"= {by / ms_dec / 1e6 / load_peaks()[0]:.2f} of the HBM peak")
ms_graph = time_graph([run], 10)
print(f"ours, replayed from a CUDA graph: {ms_graph:.3f} ms per batch of {n} images = {n / ms_graph * 1e3:.0f} images/s")
print(f"ours: {ms:.3f} ms per batch of {n} images = {n / ms * 1e3:.0f} images/s (decode 5 levels + top-k + per-level NMS + pack)")
# CPU oracle on 2 images
cells = [O.cell_anchors([sz], [0.5, 1.0, 2.0]) for sz in (32, 64, 128, 256, 512)]
anchors = O.grid_anchors([(448 // s, 448 // s) for s in strides], strides, cells, 0.0)
t0 = time.perf_counter()
props, lgs = [], []
for l in range(5):
    lg, dl = O.head_to_hwa(obj[l][:2], dlt[l][:2])
    props.append(torch.stack([O.apply_deltas(dl[i], anchors[l]) for i in range(2)])); lgs.append(lg)
want = O.find_top_rpn_proposals(props, lgs, [(448, 448)] * 2, 0.7, PRE, POST, 0.0, False)
dt = time.perf_counter() - t0
print(f"CPU oracle: {dt / 2 * 1e3:.1f} ms per image ({torch.get_num_threads()} threads)")
ob, os_, cnt, flag = out
for i in range(2):
    k = int(cnt[i]); wb, ws = want[i]
    print(f"image {i}: count {k} vs {wb.shape[0]}, max |box diff| {float((ob[i, :k].cpu() - wb).abs().max()):.2e}, logits equal {torch.equal(os_[i, :k].cpu(), ws)}")
