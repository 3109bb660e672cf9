"""Driver: the dense head as a detector runs it (BASELINE configs[3]) -- det_dense_detect = streaming select kernel +
one-CTA-per-image NMS kernel.  usage: prof_dense_detect.py [N ...] [--once] (--once: a single call per case, for ncu)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))
import torch
import det_b200 as det
from bench import time_graph, load_peaks

once = "--once" in sys.argv
ns = [int(a) for a in sys.argv[1:] if a.isdigit()] or [32, 256]
C, R = 80, 25200
strides = [8, 16, 32]
wh = [[[10, 13], [16, 30], [33, 23]], [[30, 61], [62, 45], [59, 119]], [[116, 90], [156, 198], [373, 326]]]
dh = det.DenseAnchorHead(strides, wh, C)
peak, _ = load_peaks()
dev = torch.device("cuda")
torch.zeros(1, device=dev)
for n in ns:
    g = torch.Generator(device=dev).manual_seed(3)
    pool = 1 if once else (2 if n > 64 else 4)
    heads = []
    for _ in range(pool):
        hs = [torch.randn(n, 3 * (5 + C), 640 // s, 640 // s, device=dev, generator=g) for s in strides]
        for h in hs:
            h.view(n, 3, 5 + C, h.shape[2], h.shape[3])[:, :, 4] -= 4.0
        heads.append(hs)
    for thr, cap in ((0.1, 2048), (0.25, 1024)):
        for gate in (False, True):
            ws = det.DenseDetectWorkspace(n, cap, dev)
            r = dh.detect_thresholded(heads[0], thr, 0.5, max_det=300, cand_cap=cap, gate=gate, check=True, workspace=ws)
            if once:
                torch.cuda.synchronize()
                continue
            fns = [(lambda hs=hs_: dh.detect_thresholded(hs, thr, 0.5, max_det=300, cand_cap=cap, gate=gate, check=False,
                                                         out=r, workspace=ws)) for hs_ in heads]
            ms = time_graph(fns, 10)
            k = float(r["count"].float().mean())
            cand = None
            by = n * (4 * R * (5 + C) + 36 * k + 4)
            print(json.dumps({"n": n, "thr": thr, "cap": cap, "gate": gate, "ms": round(ms, 5), "kept_per_image": k,
                              "images_per_s": round(n / ms * 1e3), "alg_GBs": round(by / ms / 1e6, 1),
                              "frac": round(by / ms / 1e6 / peak, 4)}), flush=True)
    # the unfused path for comparison: dense decode + NMS over candidates compacted by torch
    if not once:
        fn = [lambda hs=hs_: dh.decode(hs) for hs_ in heads]
        print(json.dumps({"n": n, "decode_only_ms": round(time_graph(fn, 10), 5)}), flush=True)
    del heads
    torch.cuda.empty_cache()
