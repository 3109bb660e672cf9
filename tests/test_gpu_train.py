"""GPU parity: IoU target assignment (Matcher), device subsampling, fused RPN / YOLO loss forward+backward."""
import pytest
import torch

from tests.util import gen, rand_boxes

pytestmark = pytest.mark.gpu

STRIDES = [4, 8, 16, 32, 64]
SIZES = [[32], [64], [128], [256], [512]]
RATIOS = [[0.5, 1.0, 2.0]]


@pytest.fixture(scope="module")
def det():
    import det_b200
    return det_b200


@pytest.fixture(scope="module")
def O():
    from oracle import ref_torch
    return ref_torch


def _anchors(O, img=448):
    cells = [O.cell_anchors(s, RATIOS[0]) for s in SIZES]
    return torch.cat(O.grid_anchors([(img // s, img // s) for s in STRIDES], STRIDES, cells, 0.0), 0)


def _gts(n, g, frame=448.0, gmax=16):
    out = []
    for i in range(n):
        k = int(torch.randint(1, gmax + 1, (1,), generator=g))
        b = rand_boxes(k, frame, g)
        b[:, 2:].clamp_(max=frame)
        out.append(b)
    return out


@pytest.mark.parametrize("thr,lab,lq", [((0.3, 0.7), (0, -1, 1), True), ((0.5,), (0, 1), False), ((0.3, 0.7), (0, -1, 1), False)])
def test_matcher_on_matrix_bit_exact(det, O, thr, lab, lq):
    g = gen(3)
    gt, anc = rand_boxes(23, 448.0, g), _anchors(O)[::5].contiguous()
    q = O.pairwise_iou(gt, anc)
    q[4] = 0.0  # a gt overlapping nothing: the zero-IoU low-quality quirk (matcher.py:114-120)
    q[:, 100] = q[0, 100]  # column ties -> lowest gt index
    wi, wl = O.match(q, list(thr), list(lab), lq)
    gi, gl = det.Matcher(list(thr), list(lab), lq)(q.cuda())
    assert gi.dtype == torch.int64 and gl.dtype == torch.int8
    assert torch.equal(gi.cpu(), wi) and torch.equal(gl.cpu(), wl)


def test_matcher_edges(det, O):
    m = det.Matcher([0.3, 0.7], [0, -1, 1], True)
    i, l = m(torch.zeros(0, 9).cuda())
    assert i.tolist() == [0] * 9 and l.tolist() == [0] * 9 and l.dtype == torch.int8
    q = torch.tensor([[0.3, 0.7, 0.29999998, 0.6999999, 0.0, 1.0]])
    wi, wl = O.match(q, [0.3, 0.7], [0, -1, 1], False)
    gi, gl = det.Matcher([0.3, 0.7], [0, -1, 1], False)(q.cuda())
    assert torch.equal(gl.cpu(), wl) and wl.tolist() == [-1, 1, 0, -1, 0, 1]
    with pytest.raises(AssertionError):
        m(torch.tensor([[0.5, -0.1]]).cuda())


@pytest.mark.parametrize("n", [1, 7])
def test_fused_assignment_matches_oracle(det, O, n):
    """label + matched index of every anchor bit-exact vs pairwise_iou -> Matcher per image (pre-subsample)."""
    g = gen(40 + n)
    anc = _anchors(O)
    gts = _gts(n, g)
    if n > 2:
        gts[1] = torch.zeros(0, 4)                       # image without gt
        gts[2] = torch.cat([gts[2], torch.tensor([[440.0, 440.0, 440.5, 440.5]])])  # tiny gt
    wl, wi = O.label_anchors(anc, gts)
    rpn = det.RegionProposalNetwork(STRIDES, SIZES, RATIOS)
    asg = rpn.assign(anc.cuda(), [b.cuda() for b in gts], sample=False)
    assert asg.labels.shape == (n, 50127)
    for i in range(n):
        assert torch.equal(asg.labels[i].cpu(), wl[i]), i
        assert torch.equal(asg.matched[i].cpu(), wi[i]), i


def test_device_subsample_counts_and_uniformity(det, O):
    g = gen(8)
    n, r = 6, 50127
    lab = torch.full((n, r), -1, dtype=torch.int8)
    npos = [0, 5, 128, 129, 3000, 50]
    nneg = [10, 100, 40000, 300, 47000, 0]
    for i in range(n):
        perm = torch.randperm(r, generator=g)
        lab[i, perm[:npos[i]]] = 1
        lab[i, perm[npos[i]:npos[i] + nneg[i]]] = 0
    out = det.subsample_labels_(lab.cuda().clone(), 256, 0.5, seed=1).cpu()
    for i in range(n):
        wp, wn = O.subsample_counts(npos[i], nneg[i], 256, 0.5)
        assert int((out[i] == 1).sum()) == wp and int((out[i] == 0).sum()) == wn
        assert bool(((out[i] == 1) <= (lab[i] == 1)).all()) and bool(((out[i] == 0) <= (lab[i] == 0)).all())
    # different seeds choose different subsets; the same seed is reproducible
    a = det.subsample_labels_(lab.cuda().clone(), 256, 0.5, seed=2).cpu()
    b = det.subsample_labels_(lab.cuda().clone(), 256, 0.5, seed=2).cpu()
    assert torch.equal(a, b) and not torch.equal(a, out)
    # uniformity: over many seeds every negative of image 3 (300 negatives, 128 kept) is picked ~ 128/300 of the time
    hits = torch.zeros(r)
    row = lab[3:4].cuda()
    for s in range(200):
        hits += (det.subsample_labels_(row.clone(), 256, 0.5, seed=100 + s)[0].cpu() == 0).float()
    frac = hits[lab[3] == 0] / 200.0
    assert abs(float(frac.mean()) - 128 / 300) < 1e-6 and float(frac.min()) > 0.25 and float(frac.max()) < 0.6


@pytest.mark.parametrize("num_samples,frac", [(100, 0.29), (100, 0.57), (256, 0.5), (512, 0.25), (7, 0.3)])
def test_device_subsample_positive_cap_is_the_python_double_product(det, O, num_samples, frac):
    """utils.py:63 int(num_samples * positive_fraction) in Python double: 100 * 0.29 -> 28 (fp32 would give 29)."""
    g = gen(21)
    r = 4000
    lab = torch.full((1, r), -1, dtype=torch.int8)
    perm = torch.randperm(r, generator=g)
    lab[0, perm[:600]] = 1
    lab[0, perm[600:3000]] = 0
    out = det.subsample_labels_(lab.cuda().clone(), num_samples, frac, seed=3).cpu()
    wp, wn = O.subsample_counts(600, 2400, num_samples, frac)
    assert wp == int(num_samples * frac)
    assert int((out[0] == 1).sum()) == wp and int((out[0] == 0).sum()) == wn


def test_device_subsample_dense_class_walks_a_permutation(det, O):
    """A class that holds >= 1/8 of the anchors (RPN background) is sampled by walking a pseudo-random permutation of the
    row: exact counts, subset of the class, reproducible, and uniform over the class."""
    g = gen(9)
    n, r = 4, 2048
    lab = torch.full((n, r), -1, dtype=torch.int8)
    npos, nneg = [10, 700, 0, 300], [1500, 1300, 2048, 256]
    for i in range(n):
        perm = torch.randperm(r, generator=g)
        lab[i, perm[:npos[i]]] = 1
        lab[i, perm[npos[i]:npos[i] + nneg[i]]] = 0
    out = det.subsample_labels_(lab.cuda().clone(), 256, 0.5, seed=5).cpu()
    for i in range(n):
        wp, wn = O.subsample_counts(npos[i], nneg[i], 256, 0.5)
        assert int((out[i] == 1).sum()) == wp and int((out[i] == 0).sum()) == wn, i
        assert bool(((out[i] == 1) <= (lab[i] == 1)).all()) and bool(((out[i] == 0) <= (lab[i] == 0)).all())
    again = det.subsample_labels_(lab.cuda().clone(), 256, 0.5, seed=5).cpu()
    assert torch.equal(out, again)
    hits = torch.zeros(r)
    row = lab[0:1].cuda()  # 1500 negatives, 246 kept
    trials = 300
    for s in range(trials):
        hits += (det.subsample_labels_(row.clone(), 256, 0.5, seed=1000 + s)[0].cpu() == 0).float()
    frac = hits[lab[0] == 0] / trials
    p = 246 / 1500
    assert abs(float(frac.mean()) - p) < 1e-6
    assert float(frac.min()) > p - 0.1 and float(frac.max()) < p + 0.1  # +-4.7 sigma of a binomial(300, p) frequency
    # pairs of anchors are not tied together: neighbouring negatives are kept together ~ p^2 of the time
    neg = torch.nonzero(lab[0] == 0, as_tuple=True)[0]
    both = 0
    for s in range(trials):
        o = det.subsample_labels_(row.clone(), 256, 0.5, seed=5000 + s)[0].cpu()
        both += int(((o[neg[:-1]] == 0) & (o[neg[1:]] == 0)).sum())
    assert abs(both / (trials * (neg.numel() - 1)) - p * p) < 0.01


def test_peer_sums_protocol_on_one_gpu(det):
    """det_peer_sums_publish / collect (csrc/peer.cu) with two VIRTUAL ranks whose "peer" buffers live on one GPU: every
    rank's record lands in every buffer, the sum is taken in rank order, stamps gate the read, a missing peer times out
    with NaN + error flag instead of hanging."""
    import det_b200._native as N
    dev = torch.device("cuda")
    world, slots, width = 2, 8, 8
    bufs = [torch.zeros(slots * world * 16, device=dev) for _ in range(world)]
    peers = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    g = gen(4)
    for step in range(1, 20):
        vecs = [torch.randn(width, generator=g).to(dev) for _ in range(world)]
        for r in range(world):
            N.call("det_peer_sums_publish", N.ptr(vecs[r]), width, r, world, N.ptr(peers), slots, step % slots, step,
                   N.stream())
        for r in range(world):
            out = torch.empty(width, device=dev)
            N.call("det_peer_sums_collect", N.ptr(out), width, world, N.ptr(bufs[r]), slots, step % slots, step,
                   10 ** 9, N.ptr(err), N.stream())
            assert torch.equal(out, vecs[0] + vecs[1])  # rank order, fp32: bitwise the same on every rank
    assert int(err.item()) == 0
    # rank 1 never publishes step 20: the collect gives up after 20 ms
    N.call("det_peer_sums_publish", N.ptr(vecs[0]), width, 0, world, N.ptr(peers), slots, 20 % slots, 20, N.stream())
    out = torch.zeros(width, device=dev)
    N.call("det_peer_sums_collect", N.ptr(out), width, world, N.ptr(bufs[0]), slots, 20 % slots, 20, 20_000_000,
           N.ptr(err), N.stream())
    assert int(err.item()) == 1 and bool(torch.isnan(out).all())
    # the single-process PeerSums wrapper (world 1) returns what was published, one step late
    ps = det.dist.PeerSums(dev)
    a, b = torch.arange(8.0, device=dev), torch.ones(8, device=dev)
    ps.publish(a); ps.publish(b)
    assert torch.equal(ps.collect(), a) and torch.equal(ps.collect(), b)
    ps.check()
    px = det.dist.PeerSums(dev)  # one launch per step: publish(t) + collect(t - 1)
    assert px.exchange(a) is None
    assert torch.equal(px.exchange(b), a) and torch.equal(px.exchange(a + b), b) and torch.equal(px.flush(), a + b)
    px.check()


def test_yolo_loss_publishes_from_its_last_cta(det):
    """det_yolo_loss_peer (world 1): same losses and gradients as det_yolo_loss, and "world_sums_prev" of step t is the
    sums vector of step t - 1 -- published and collected by the loss kernel's own last CTA."""
    dev = torch.device("cuda")
    yh = det.YoloGridHead(7, 2, 20, (448, 448))
    tr = det.YoloGridTrainer(yh)
    g = gen(12)
    n = 300
    gts = [torch.cat([xy, xy + wh], 1) for xy, wh in
           ((torch.rand(k, 2, generator=g) * 300, torch.rand(k, 2, generator=g) * 100 + 8)
            for k in torch.randint(1, 6, (n,), generator=g).tolist())]
    gtc = torch.cat([torch.randint(0, 20, (b.shape[0],), generator=g) for b in gts]).to(dev)
    off = torch.tensor([0] + [b.shape[0] for b in gts]).cumsum(0).to(torch.int32).to(dev)
    gtb = torch.cat(gts).to(dev)
    asg = tr.assign_packed(gtb, off, n)
    ps = det.dist.PeerSums(dev)
    prev = None
    for step in range(6):
        head = torch.randn(n, 7, 7, 30, generator=g).to(dev)
        want = tr.loss(head, asg, gtc, with_grads=True)
        got = tr.loss(head, asg, gtc, with_grads=True, peer=ps)
        assert torch.equal(got["grad_head"], want["grad_head"])
        torch.testing.assert_close(got["sums"], want["sums"], rtol=1e-5, atol=1e-6)  # atomics: order of the CTAs
        if step == 0:
            assert got["world_sums_prev"] is None
        else:
            assert torch.equal(got["world_sums_prev"], prev)
        prev = got["sums"].clone()
    ps.check()


def test_fused_exchange_replays_from_a_cuda_graph(det):
    """PeerSums(graph_safe=True): the step stamp lives on the device, so a captured training step can be replayed; after
    every replay the collected vector is the sums of the previous replay."""
    dev = torch.device("cuda")
    yh = det.YoloGridHead(7, 2, 20, (448, 448))
    tr = det.YoloGridTrainer(yh)
    g = gen(13)
    n = 64
    gts = [torch.tensor([[10.0 + i, 20.0, 120.0 + i, 200.0]]) for i in range(n)]
    gtc = torch.randint(0, 20, (n,), generator=g).to(dev)
    off = torch.arange(n + 1, dtype=torch.int32).to(dev)
    gtb = torch.cat(gts).to(dev)
    head = torch.randn(n, 7, 7, 30, generator=g).to(dev)
    ps = det.dist.PeerSums(dev, graph_safe=True)
    res = {}
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        res = tr.loss(head, tr.assign_packed(gtb, off, n), gtc, with_grads=True, peer=ps)  # stamp 1 (eager warm-up)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        res = tr.loss(head, tr.assign_packed(gtb, off, n), gtc, with_grads=True, peer=ps)
    want = tr.loss(head, tr.assign_packed(gtb, off, n), gtc, with_grads=True)["sums"]
    for rep in range(4):
        head.mul_(1.0 + 0.1 * rep)  # new inputs in place: the graph reads the same buffers
        expect_prev = want.clone()
        want = tr.loss(head, tr.assign_packed(gtb, off, n), gtc, with_grads=True)["sums"].clone()
        graph.replay()
        torch.cuda.synchronize()
        torch.testing.assert_close(res["sums"], want, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(res["world_sums_prev"], expect_prev, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ps.flush(), want, rtol=1e-5, atol=1e-6)  # the published vector is the scaled one
    ps.check()


def _rpn_loss_case(O, n, seed, beta=0.0):
    g = gen(seed)
    anc = _anchors(O)
    gts = _gts(n, g)
    labs, idxs = O.label_anchors(anc, gts)
    labs = [O.rpn_subsample_(l, 256, 0.5) for l in labs]
    matched_boxes = [gt[i] for gt, i in zip(gts, idxs)]
    logits = torch.randn(n, anc.shape[0], generator=g)
    deltas = torch.randn(n, anc.shape[0], 4, generator=g) * 0.5
    return anc, gts, labs, idxs, matched_boxes, logits, deltas


@pytest.mark.parametrize("beta", [0.0, 0.11])
def test_rpn_loss_forward_backward_vs_oracle(det, O, beta):
    n = 4
    torch.manual_seed(0)
    anc, gts, labs, idxs, mboxes, logits, deltas = _rpn_loss_case(O, n, 5, beta)
    lg = logits.clone().requires_grad_(True)
    dl = deltas.clone().requires_grad_(True)
    want = O.rpn_losses(anc, lg, torch.stack(labs), dl, torch.stack(mboxes), smooth_l1_beta=beta)
    (want["cls_loss"] + 2.0 * want["loc_loss"]).backward()
    rpn = det.RegionProposalNetwork(STRIDES, SIZES, RATIOS, smooth_l1_beta=beta)
    # (a) the batched fast path with autograd
    table = torch.cat(gts, 0).cuda()
    off = torch.tensor([0] + list(torch.tensor([len(x) for x in gts]).cumsum(0)), dtype=torch.int32).cuda()
    asg = det.Assignment(torch.stack(labs).cuda(), torch.stack(idxs).cuda(), table, off)
    glg = logits.cuda().requires_grad_(True)
    gdl = deltas.cuda().requires_grad_(True)
    res = rpn.fused_losses(anc.cuda(), glg, gdl, asg)
    (res["cls_loss"] + 2.0 * res["loc_loss"]).backward()
    torch.testing.assert_close(res["cls_loss"].cpu(), want["cls_loss"].detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(res["loc_loss"].cpu(), want["loc_loss"].detach(), rtol=1e-5, atol=1e-7)
    assert int(res["num_pos_anchors"]) == want["num_pos"] and int(res["num_neg_anchors"]) == want["num_neg"]
    torch.testing.assert_close(glg.grad.cpu(), lg.grad, rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(gdl.grad.cpu(), dl.grad, rtol=1e-5, atol=1e-9)
    # (b) single-launch fused forward+backward (upstream gradient 1 for both terms)
    res2 = rpn.fused_losses(anc.cuda(), logits.cuda(), deltas.cuda(), asg, with_grads=True)
    lg2 = logits.clone().requires_grad_(True)
    dl2 = deltas.clone().requires_grad_(True)
    w2 = O.rpn_losses(anc, lg2, torch.stack(labs), dl2, torch.stack(mboxes), smooth_l1_beta=beta)
    (w2["cls_loss"] + w2["loc_loss"]).backward()
    torch.testing.assert_close(res2["grad_logits"].cpu(), lg2.grad, rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(res2["grad_deltas"].cpu(), dl2.grad, rtol=1e-5, atol=1e-9)
    # (c) the reference's own signature: lists per level / per image
    lsz = [37632, 9408, 2352, 588, 147]
    out = rpn.losses([det.Boxes(a.cuda()) for a in torch.split(anc, lsz)],
                     [x.cuda() for x in torch.split(logits, lsz, dim=1)], [l.cuda() for l in labs],
                     [x.cuda() for x in torch.split(deltas, lsz, dim=1)], [m.cuda() for m in mboxes])
    torch.testing.assert_close(out["cls_loss"].cpu(), want["cls_loss"].detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(out["loc_loss"].cpu(), want["loc_loss"].detach(), rtol=1e-5, atol=1e-7)


def test_label_and_sample_anchors_reference_signature(det, O):
    g = gen(12)
    anc = _anchors(O)
    gts = _gts(3, g)
    rpn = det.RegionProposalNetwork(STRIDES, SIZES, RATIOS)
    lsz = [37632, 9408, 2352, 588, 147]
    insts = []
    for b in gts:
        inst = det.Instances((448, 448))
        inst.gt_boxes = det.Boxes(b.cuda())
        insts.append(inst)
    labels, mboxes = rpn.label_and_sample_anchors([det.Boxes(a.cuda()) for a in torch.split(anc, lsz)], insts)
    wl, wi = O.label_anchors(anc, gts)
    for i in range(3):
        l = labels[i].cpu()
        npos_avail, nneg_avail = int((wl[i] == 1).sum()), int((wl[i] == 0).sum())
        wp, wn = O.subsample_counts(npos_avail, nneg_avail, 256, 0.5)
        assert int((l == 1).sum()) == wp and int((l == 0).sum()) == wn
        assert bool(((l == 1) <= (wl[i] == 1)).all()) and bool(((l == 0) <= (wl[i] == 0)).all())
        assert torch.equal(mboxes[i].cpu(), gts[i][wi[i]])


def test_yolo_assign_and_loss_vs_oracle(det, O):
    g = gen(31)
    n = 6
    yh = det.YoloGridHead(7, 2, 20, (448, 448))
    tr = det.YoloGridTrainer(yh)
    gts = _gts(n, g)
    gcls = [torch.randint(0, 20, (len(b),), generator=g) for b in gts]
    head = torch.randn(n, 7, 7, 30, generator=g)
    anchors = O.yolo_grid_anchors(7, (448, 448), yh.priors)
    assert torch.equal(tr.prior_boxes("cuda").cpu(), anchors)
    wl, wi = O.label_anchors(anchors, gts)
    asg = tr.assign([b.cuda() for b in gts])
    assert torch.equal(asg.labels.cpu(), torch.stack(wl)) and torch.equal(asg.matched.cpu(), torch.stack(wi))
    hc = head.clone().requires_grad_(True)
    want = O.yolo_loss(hc, torch.stack(wl), torch.stack(wi), gts, gcls, 2, 20, (448, 448), yh.priors)
    (want["loc_loss"] + want["obj_loss"] + want["cls_loss"]).backward()
    hg = head.cuda().requires_grad_(True)
    res = tr.loss(hg, asg, torch.cat(gcls).cuda())
    (res["loc_loss"] + res["obj_loss"] + res["cls_loss"]).backward()
    for k in ("loc_loss", "obj_loss", "cls_loss"):
        torch.testing.assert_close(res[k].cpu(), want[k].detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(hg.grad.cpu(), hc.grad, rtol=1e-5, atol=1e-8)
    res2 = tr.loss(head.cuda(), asg, torch.cat(gcls).cuda(), with_grads=True)
    torch.testing.assert_close(res2["grad_head"].cpu(), hc.grad, rtol=1e-5, atol=1e-8)


def test_forward_training_from_heads(det, O):
    g = gen(19)
    n, img = 2, 224
    obj = [torch.randn(n, 3, img // s, img // s, generator=g) for s in STRIDES]
    dlt = [torch.randn(n, 12, img // s, img // s, generator=g) * 0.3 for s in STRIDES]
    gts = _gts(n, g, 224.0)
    insts = []
    for b in gts:
        inst = det.Instances((img, img))
        inst.gt_boxes = det.Boxes(b.cuda())
        insts.append(inst)
    rpn = det.RegionProposalNetwork(STRIDES, SIZES, RATIOS).train()
    o_g = [o.cuda().requires_grad_(True) for o in obj]
    d_g = [d.cuda().requires_grad_(True) for d in dlt]
    props, losses = rpn.forward([(img, img)] * n, head_outputs=(o_g, d_g), gt_instances=insts)
    assert set(losses) == {"cls_loss", "loc_loss"} and len(props) == n
    (losses["cls_loss"] + losses["loc_loss"]).backward()
    assert all(t.grad is not None and torch.isfinite(t.grad).all() for t in o_g + d_g)
    assert float(losses["cls_loss"]) > 0 and all(len(p) <= 1000 for p in props)


def test_rpn_giou_loss_forward_backward_vs_oracle(det, O):
    """box_reg_loss_type="giou" (reference box_regression.py:159-165; fvcore giou_loss restated by the oracle --
    third-party arithmetic, unpinned by the reference): value and gradients against autograd through the oracle."""
    n = 3
    torch.manual_seed(0)
    anc, gts, labs, idxs, mboxes, logits, deltas = _rpn_loss_case(O, n, 9, 0.0)
    deltas = deltas * 0.6
    deltas[0, :50, 2] = 6.0  # beyond scale_clamp: the clamp must stop the gradient
    lg = logits.clone().requires_grad_(True)
    dl = deltas.clone().requires_grad_(True)
    want = O.rpn_losses(anc, lg, torch.stack(labs), dl, torch.stack(mboxes), box_reg_loss_type="giou")
    (want["cls_loss"] + 3.0 * want["loc_loss"]).backward()
    rpn = det.RegionProposalNetwork(STRIDES, SIZES, RATIOS, box_reg_loss_type="giou")
    table = torch.cat(gts, 0).cuda()
    off = torch.tensor([0] + list(torch.tensor([len(x) for x in gts]).cumsum(0)), dtype=torch.int32).cuda()
    labs_t = torch.stack(labs)
    labs_t[0, :50] = 1  # make the clamped rows foreground (matched gt index stays valid)
    want = O.rpn_losses(anc, lg, labs_t, dl, torch.stack(mboxes), box_reg_loss_type="giou")
    lg.grad = None
    dl.grad = None
    (want["cls_loss"] + 3.0 * want["loc_loss"]).backward()
    asg = det.Assignment(labs_t.cuda(), torch.stack(idxs).cuda(), table, off)
    glg = logits.cuda().requires_grad_(True)
    gdl = deltas.cuda().requires_grad_(True)
    res = rpn.fused_losses(anc.cuda(), glg, gdl, asg)
    (res["cls_loss"] + 3.0 * res["loc_loss"]).backward()
    torch.testing.assert_close(res["loc_loss"].cpu(), want["loc_loss"].detach(), rtol=2e-5, atol=1e-7)
    torch.testing.assert_close(res["cls_loss"].cpu(), want["cls_loss"].detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(gdl.grad.cpu(), dl.grad, rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(glg.grad.cpu(), lg.grad, rtol=1e-5, atol=1e-9)
    assert float(gdl.grad[0, :50, 2].abs().max()) == 0.0
    with pytest.raises(ValueError):
        det.RegionProposalNetwork(STRIDES, SIZES, RATIOS, box_reg_loss_type="l2").fused_losses(
            anc.cuda(), logits.cuda(), deltas.cuda(), asg)
