"""Top-k NMS (max_out << boxes) on the BASELINE configs[4] shapes: 8 images x M boxes, IoU 0.5, first 1000 kept.
Compares det_nms_batched's top-k tier with the same call asked for every kept index (no tier) and with torchvision CUDA."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))
import torch, torchvision
import det_b200 as det

def t(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it

rows = []
for M in (10000, 25200, 50000, 100000):
    for ncat in (80, 1):
        g = torch.Generator().manual_seed(4)
        xy = torch.rand(8, M, 2, generator=g) * 0.8 * 1024; wh = torch.rand(8, M, 2, generator=g) * 0.2 * 1024 + 1
        b = torch.cat([xy, xy + wh], 2).cuda()
        s = torch.stack([(torch.randperm(M, generator=g).float() + 0.5) / M for _ in range(8)]).cuda()
        c = torch.randint(0, ncat, (8, M), generator=g).cuda()
        top, tc = det.nms_images(b, s, c, None, 0.5, 1000)
        full, fc = det.nms_images(b, s, c, None, 0.5, None)
        same = all(torch.equal(top[i, :int(tc[i])], full[i, :int(tc[i])]) for i in range(8))
        tv = [torchvision.ops.batched_nms(b[i], s[i], c[i], 0.5)[:1000] for i in range(8)]
        same_tv = all(torch.equal(top[i, :int(tc[i])], tv[i]) for i in range(8))
        ms_top = t(lambda: det.nms_images(b, s, c, None, 0.5, 1000))
        ms_full = t(lambda: det.nms_images(b, s, c, None, 0.5, None), 3)
        ms_tv = t(lambda: [torchvision.ops.batched_nms(b[i], s[i], c[i], 0.5)[:1000] for i in range(8)], 2)
        rows.append({"M": M, "categories": ncat, "ms_top1000": round(ms_top, 4), "ms_all_kept": round(ms_full, 4),
                     "ms_torchvision_cuda_loop": round(ms_tv, 3), "top1000_equals_prefix_of_full": same,
                     "equals_torchvision_cuda": same_tv, "kept_top": int(tc.min())})
        print(json.dumps(rows[-1]), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "sweep_topk.json"), "w"), indent=1)
