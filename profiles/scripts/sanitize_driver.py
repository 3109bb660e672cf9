"""One small invocation of every kernel family of libdet_b200.so, for compute-sanitizer (profiles/scripts/sanitize_r02.sh).
Sizes are chosen so that each special path is taken once (cooperative huge-segment sweep, large path, tier cuts, lazy
sampler, peer exchange) while a racecheck run still finishes in minutes.  Usage: sanitize_driver.py [section ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "object-detection-pytorch-rust_b200"))
import torch  # noqa: E402
import det_b200 as det  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)


def boxes(m, frame=512.0, size=80.0):
    xy = torch.rand(m, 2, generator=g) * frame * 0.8
    wh = torch.rand(m, 2, generator=g) * size + 1
    return torch.cat([xy, xy + wh], 1)


def sec_yolo():
    yh = det.YoloGridHead(7, 2, 20, (448, 448))                      # yolo_fast_kernel<640,2,7,2,20>
    yh.detect(torch.randn(12, 7, 7, 30, generator=g).to(dev), 0.25, 0.5, max_det=300)
    yh.detect(torch.randn(4, 7, 7, 30, generator=g).to(dev), 0.05, 0.5, max_det=20)   # tiers fall through
    h = torch.randn(3, 7, 7, 30, generator=g)
    h[0, 0, 0, 0] = float("nan")                                     # slow exact path
    yh.detect(h.to(dev), 0.25, 0.5)
    y2 = det.YoloGridHead(5, 3, 7, (320, 320))                       # generic instantiation
    y2.detect(torch.randn(5, 5, 5, 22, generator=g).to(dev), 0.2, 0.5)
    y3 = det.YoloGridHead(9, 2, 10, (288, 288))                      # > 128 predictors: yolo_decode_nms_kernel
    y3.detect(torch.randn(2, 9, 9, 20, generator=g).to(dev), 0.3, 0.5)


def sec_nms():
    for m, ncat in ((300, 5), (3000, 80), (6000, 1), (9000, 3)):     # small path, CTA segments, huge (cooperative) segments
        b = boxes(m).to(dev)
        s = torch.rand(m, generator=g).to(dev)
        c = torch.randint(0, ncat, (m,), generator=g).to(dev)
        det.batched_nms(b, s, c, 0.5)
    n, m = 3, 5000                                                   # batched large path with a top-k tier
    b = torch.stack([boxes(m) for _ in range(n)]).to(dev)
    s = torch.rand(n, m, generator=g).to(dev)
    c = torch.randint(0, 80, (n, m), generator=g).to(dev)
    det.nms_images(b, s, c, None, 0.5, 200)
    det.nms_images(b, s, c, None, 0.5, m)


def sec_dense():
    strides = [8, 16, 32]
    wh = [[[10, 13], [16, 30], [33, 23]], [[30, 61], [62, 45], [59, 119]], [[116, 90], [156, 198], [373, 326]]]
    dh = det.DenseAnchorHead(strides, wh, 80)
    n = 2
    hs = [torch.randn(n, 3 * 85, 320 // s, 320 // s, generator=g).to(dev) for s in strides]
    for h in hs:
        h.view(n, 3, 85, h.shape[2], h.shape[3])[:, :, 4] -= 3.0
    dh.decode(hs)
    for gate in (False, True):
        dh.detect_thresholded(hs, 0.1, 0.5, max_det=100, cand_cap=2048, gate=gate, check=True)


def sec_train():
    strides = (4, 8, 16, 32, 64)
    rpn = det.RegionProposalNetwork(list(strides))
    hw = [(192 // s, 192 // s) for s in strides]
    anchors = torch.cat(rpn.anchor_generator.grid_anchors(hw, dev), 0)
    grid = rpn.anchor_generator.grid_layout(hw)
    n = 6
    gts = [boxes(int(k), 192.0, 60.0).to(dev) for k in (1, 5, 0, 16, 40, 3)]
    asg = rpn.assign(anchors, gts, sample=True, seed=3)                            # match_pass1/2 + subsample_kernel
    R = anchors.shape[0]
    logits = torch.randn(n, R, generator=g).to(dev)
    deltas = (torch.randn(n, R, 4, generator=g) * 0.5).to(dev)
    rpn.fused_losses(anchors, logits, deltas, asg, with_grads=True)               # rpn_loss_kernel<0>
    rpn.box_reg_loss_type = "giou"
    rpn.fused_losses(anchors, logits, deltas, asg, with_grads=True)               # rpn_loss_kernel<1>
    rpn.box_reg_loss_type = "smooth_l1"
    asg2 = rpn.assign(anchors, gts, sample=True, seed=3, grid=grid)                # match_rowmax + match_grid + grid subsample
    obj = [torch.randn(n, 3, h, w, generator=g).to(dev) for h, w in hw]
    dlt = [(torch.randn(n, 12, h, w, generator=g) * 0.5).to(dev) for h, w in hw]
    g_obj, g_dlt = [torch.zeros_like(o) for o in obj], [torch.zeros_like(d) for d in dlt]
    rpn._run_sampled(anchors, obj, dlt, asg2, n, None, (g_obj, g_dlt), None)       # rpn_loss_sampled_kernel
    table, off = rpn.anchor_matcher.pack_gt(gts, dev)
    asg3 = rpn.assign_sampled(anchors, table, off, n, grid, seed=4)                # assign_candidates + subsample_lazy
    rpn._run_sampled(anchors, obj, dlt, asg3, n, None, (g_obj, g_dlt), (asg2.samples, asg2.sample_count))
    sizes = torch.tensor([[192, 192]] * n, dtype=torch.int32, device=dev)
    rpn.proposals_from_heads(obj, dlt, sizes)                                     # rpn decode + select + NMS tiers
    # grid head training step incl. the fused peer exchange (single rank: the protocol still runs)
    yh = det.YoloGridHead(7, 2, 20, (448, 448))
    tr = det.YoloGridTrainer(yh)
    gb = torch.cat([boxes(3, 448.0, 90.0) for _ in range(8)]).to(dev)
    gc = torch.randint(0, 20, (24,), generator=g).to(dev)
    goff = (torch.arange(9, dtype=torch.int32) * 3).to(dev)
    head = torch.randn(8, 7, 7, 30, generator=g).to(dev)
    a = tr.assign_packed(gb, goff, 8)
    res = tr.loss(head, a, gc, with_grads=True)
    ps = det.dist.PeerSums(dev)
    ps.exchange(res["sums"])
    ps.exchange(res["sums"])
    ps.flush()
    ps.check()
    ps2 = det.dist.PeerSums(dev)
    tr.loss(head, a, gc, with_grads=True, peer=ps2)                                # det_yolo_loss_peer
    tr.loss(head, a, gc, with_grads=True, peer=ps2)
    ps2.flush()
    ps2.check()


def sec_misc():
    b1, b2 = det.Boxes(boxes(700).to(dev)), det.Boxes(boxes(900).to(dev))
    det.pairwise_iou(b1, b2)
    det.pairwise_ioa(b1, b2)
    t = det.Box2BoxTransform((1.0, 1.0, 1.0, 1.0))
    d = t.get_deltas(b1.tensor, boxes(700).to(dev))
    t.apply_deltas(d, b1.tensor)


SECTIONS = {"yolo": sec_yolo, "nms": sec_nms, "dense": sec_dense, "train": sec_train, "misc": sec_misc}

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(SECTIONS)):
        SECTIONS[name]()
        torch.cuda.synchronize()
        print("section ok:", name, flush=True)
