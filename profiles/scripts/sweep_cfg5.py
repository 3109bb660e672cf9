"""BASELINE configs[4]: pairwise IoU + NMS stress sweep, M = 1k ... 100k boxes per image, 8 images per GPU.
Ours (whole batch, one call) vs the existing Blackwell kernels reached by the reference: torchvision.ops.batched_nms on
CUDA (one call per image, as python/src/models/utils.py:74-95 loops) and the reference's torch-op pairwise_iou on CUDA.
Kept indices are checked against torchvision's CUDA result (distinct scores) and, up to 20k boxes, the C oracle."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))
import torch, torchvision
import det_b200 as det
from oracle import ref_torch as O
dev = torch.device("cuda", 0)
def t(fn, it):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it
def torch_pairwise_iou(b1, b2):  # the reference's op sequence (structures/boxes.py:173-214) on CUDA tensors
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1]); a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    wh = torch.min(b1[:, None, 2:], b2[:, 2:]) - torch.max(b1[:, None, :2], b2[:, :2])
    wh.clamp_(min=0); inter = wh.prod(dim=2)
    return torch.where(inter > 0, inter / (a1[:, None] + a2 - inter), torch.zeros(1, dtype=inter.dtype, device=inter.device))
rows = []
NIMG = 8
for M in (1000, 2000, 5000, 10000, 20000, 50000, 100000):
    g = torch.Generator().manual_seed(4)
    xy = torch.rand(NIMG, M, 2, generator=g) * 0.8 * 1024; wh = torch.rand(NIMG, M, 2, generator=g) * 0.2 * 1024 + 1
    boxes = torch.cat([xy, xy + wh], 2)
    scores = torch.stack([(torch.randperm(M, generator=g).float() + 0.5) / M for _ in range(NIMG)])
    for ncat in (80, 1):
        cats = torch.randint(0, ncat, (NIMG, M), generator=g)
        b, s, c = boxes.to(dev), scores.to(dev), cats.to(dev)
        keep, cnt = det.nms_images(b, s, c, None, 0.5, M, mode=1)
        torch.cuda.synchronize()
        it = 10 if M <= 20000 else 3
        ms_ours = t(lambda: det.nms_images(b, s, c, None, 0.5, M, mode=1), it)
        def tv():
            return [torchvision.ops.batched_nms(b[i], s[i], c[i], 0.5) for i in range(NIMG)]
        ms_tv = t(tv, it)
        ref = tv()
        same_tv = all(int(cnt[i]) == ref[i].numel() and torch.equal(keep[i, :ref[i].numel()], ref[i]) for i in range(NIMG))
        same_or = None
        if M <= 20000:
            w = O.batched_nms(boxes[0], scores[0], cats[0], 0.5) if M > 1000 else None
            if w is not None:
                same_or = int(cnt[0]) == w.numel() and torch.equal(keep[0, :w.numel()].cpu(), w)
        rows.append({"M": M, "categories": ncat, "images": NIMG, "ms_ours_batch": ms_ours, "ms_torchvision_cuda_loop": ms_tv,
                     "speedup_vs_torchvision_cuda": ms_tv / ms_ours, "kept_per_image": float(cnt.float().mean()),
                     "kept_equal_torchvision_cuda": bool(same_tv), "kept_equal_c_oracle_img0": same_or})
        print(rows[-1], flush=True)
    if M <= 20000:
        b1 = boxes[0].to(dev)
        out = det.pairwise_iou(det.Boxes(b1), det.Boxes(b1)); torch.cuda.synchronize()
        ms_o = t(lambda: det.pairwise_iou(det.Boxes(b1), det.Boxes(b1)), 5)
        ms_r = t(lambda: torch_pairwise_iou(b1, b1), 5)
        want = torch_pairwise_iou(b1, b1)
        by = 32 * M + 4 * M * M
        rows.append({"M": M, "pairwise_iou_ms_ours": ms_o, "pairwise_iou_ms_torch_ops_cuda": ms_r, "algorithmic_GBs_ours": by / ms_o / 1e6,
                     "frac_hbm": by / ms_o / 1e6 / 6537.6, "equal_to_torch_ops": bool(torch.equal(out, want))})
        print(rows[-1], flush=True)
        del out, want
        torch.cuda.empty_cache()
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "sweep_cfg5.json"), "w"), indent=1)
