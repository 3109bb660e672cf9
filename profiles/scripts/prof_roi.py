"""ROIPooler forward: ours (level assignment + one kernel for the pyramid) vs the reference's per-level loop over
torchvision's CUDA roi_align (roi_poolers.py:269-331), FPN-18 features (64 ch, strides 4..32) at 448x448."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))
import torch, torchvision
import det_b200 as det
dev = torch.device("cuda", 0)
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it
for n_img, per_img, C in ((16, 512, 64), (64, 512, 64), (16, 512, 256)):
    g = torch.Generator().manual_seed(1)
    strides = [4, 8, 16, 32]; scales = [1.0 / s for s in strides]
    feats = [torch.randn(n_img, C, 448 // s, 448 // s, generator=g).to(dev) for s in strides]
    bl = []
    for i in range(n_img):
        sz = torch.exp(torch.rand(per_img, generator=g) * math.log(300.0 / 8.0)) * 8.0
        cx, cy = torch.rand(per_img, generator=g) * 448, torch.rand(per_img, generator=g) * 448
        bl.append(torch.stack((cx - sz / 2, cy - sz / 2, cx + sz / 2, cy + sz / 2), 1).to(dev))
    pooler = det.ROIPooler(scales, 0, 7, "ROIAlignV2")
    boxes = [det.Boxes(b) for b in bl]
    def ours(): return pooler(feats, boxes)
    def ref():
        rois = torch.cat([torch.cat((torch.full_like(b[:, :1], i), b), 1) for i, b in enumerate(bl)], 0)
        allb = torch.cat(bl, 0)
        area = (allb[:, 2] - allb[:, 0]) * (allb[:, 3] - allb[:, 1])
        lv = torch.clamp(torch.floor(4 + torch.log2(torch.sqrt(area) / 224 + 1e-8)), min=2, max=5).to(torch.int64) - 2
        out = torch.zeros((rois.shape[0], C, 7, 7), device=dev)
        for l, (f, s) in enumerate(zip(feats, scales)):
            inds = torch.nonzero(lv == l, as_tuple=True)[0]
            out.index_put_((inds,), torchvision.ops.roi_align(f, rois[inds], (7, 7), s, 0, True))
        return out
    a, b = ours(), ref()
    err = float((a - b).abs().max())
    mo, mr = t(ours), t(ref)
    m = n_img * per_img
    print(f"N={n_img} boxes={m} C={C}: ours {mo*1e3:7.1f} us ({m*C*49*4/mo/1e6:6.0f} GB/s of output) | reference loop on CUDA {mr*1e3:7.1f} us | x{mr/mo:.1f} | max diff {err:.1e}")
