"""GPU parity: the grid-anchor training path -- det_match_grid (one-pass assignment), det_subsample_labels_grid
(O(samples) subsample) and det_rpn_loss_sampled (loss forward + backward on the NCHW head) -- against the CPU oracle and
against the generic kernels they replace (which are themselves oracle-checked in test_gpu_train.py)."""
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


def _pyramid(det, img, strides=STRIDES, sizes=SIZES, ratios=RATIOS):
    rpn = det.RegionProposalNetwork(strides, sizes, ratios)
    hw = [(img // s, img // s) for s in strides]
    levels = rpn.anchor_generator.grid_anchors(hw, torch.device("cuda"))
    return rpn, hw, levels, torch.cat(levels, 0), rpn.anchor_generator.grid_layout(hw)


def _gts(n, g, frame=448.0, lo=1, hi=17):
    out = []
    for _ in range(n):
        k = int(torch.randint(lo, hi, (1,), generator=g))
        xy = torch.rand(k, 2, generator=g) * 0.8 * frame
        wh = torch.rand(k, 2, generator=g) * 0.2 * frame + 1
        out.append(torch.cat([xy, (xy + wh).clamp(max=frame)], 1))
    return out


def _same(a, b):
    return torch.equal(a, b)


@pytest.mark.parametrize("img,n", [(448, 9), (224, 3), (200, 4)])  # 200: level sizes 50/25/12/6/3 -- ragged tiles
def test_grid_matcher_equals_oracle_and_generic(det, O, img, n):
    g = gen(100 + img + n)
    rpn, hw, levels, at, grid = _pyramid(det, img)
    gts = _gts(n, g, float(img))
    if n > 3:
        gts[1] = torch.zeros(0, 4)                                                       # image without gt
        gts[2] = torch.cat([gts[2], torch.tensor([[440.0, 440.0, 440.5, 440.5]])])       # tiny gt (low-quality only)
        gts[3] = torch.cat([gts[3], torch.tensor([[-500.0, -500.0, -400.0, -450.0]])])   # gt off the image: row max 0
    if n > 8:
        gts[4] = torch.cat([gts[4], torch.tensor([[50.0, 60.0, 50.0, 90.0]])])           # zero-area gt: row max 0
        gts[5] = rand_boxes(70, float(img), g)                                           # more than 32 gts
        gts[6] = torch.cat([gts[6], gts[6][:2]])                                         # duplicated gts: argmax ties
        gts[7] = torch.cat([gts[7], torch.tensor([[float("nan"), 5.0, 60.0, 70.0]])])    # NaN gt: IoU 0 everywhere
    dev_gts = [b.cuda() for b in gts]
    wl, wi = O.label_anchors(at.cpu(), gts)
    m = rpn.anchor_matcher
    got_i, got_l, got_q, _, _ = m.match_boxes(dev_gts, at, return_iou=True, grid=grid)
    gen_i, gen_l, gen_q, _, _ = m.match_boxes(dev_gts, at, return_iou=True)
    assert _same(got_l, gen_l) and _same(got_i, gen_i) and _same(got_q, gen_q)
    for i in range(n):
        assert torch.equal(got_l[i].cpu(), wl[i]), i
        assert torch.equal(got_i[i].cpu(), wi[i]), i


@pytest.mark.parametrize("thresholds,labels,lq", [([0.5], [0, 1], False), ([0.3, 0.7], [0, -1, 1], False),
                                                   ([0.2, 0.4, 0.6], [0, -1, 0, 1], True)])
def test_grid_matcher_other_rules(det, O, thresholds, labels, lq):
    g = gen(7)
    rpn, hw, levels, at, grid = _pyramid(det, 320)
    gts = [b.cuda() for b in _gts(5, g, 320.0)]
    m = det.Matcher(thresholds, labels, lq)
    a = m.match_boxes(gts, at, grid=grid)
    b = m.match_boxes(gts, at)
    assert _same(a[0], b[0]) and _same(a[1], b[1])
    for i, gt in enumerate(gts):
        q = O.pairwise_iou(gt.cpu(), at.cpu())
        wi, wl = O.match(q, thresholds, labels, lq)
        assert torch.equal(a[0][i].cpu(), wi) and torch.equal(a[1][i].cpu(), wl)


def test_grid_matcher_single_level_one_anchor_and_nine(det, O):
    g = gen(3)
    for sizes, ratios in (([[64]], [[1.0]]), ([[32, 64, 128]], [[0.5, 1.0, 2.0]])):
        rpn, hw, levels, at, grid = _pyramid(det, 256, [16], sizes, ratios)
        assert grid[1] in (1, 9)
        gts = [b.cuda() for b in _gts(4, g, 256.0)]
        a = rpn.anchor_matcher.match_boxes(gts, at, grid=grid)
        b = rpn.anchor_matcher.match_boxes(gts, at)
        assert _same(a[0], b[0]) and _same(a[1], b[1])
        wl, wi = O.label_anchors(at.cpu(), [x.cpu() for x in gts])
        assert torch.equal(a[1].cpu(), torch.stack(wl)) and torch.equal(a[0].cpu(), torch.stack(wi))


def test_grid_matcher_stats_are_exact(det):
    g = gen(5)
    rpn, hw, levels, at, grid = _pyramid(det, 448)
    gts = [b.cuda() for b in _gts(6, g)]
    matched, labels, stats, table, off = rpn.anchor_matcher.match_boxes(gts, at, grid=grid, with_stats=True)
    cnt = stats.counts.cpu()
    for i in range(6):
        pos = (labels[i] == 1).nonzero()[:, 0]
        assert int(cnt[i, 0]) == pos.numel() and int(cnt[i, 1]) == int((labels[i] == -1).sum())
        lst = stats.pos_list[i, :pos.numel()].cpu()
        assert bool(((lst >> 24) == 1).all())
        assert torch.equal((lst & 0xffffff).sort().values, pos.cpu().to(torch.int32))


@pytest.mark.parametrize("num_samples,frac", [(256, 0.5), (64, 0.25), (512, 0.5)])
def test_grid_subsample_is_bit_identical_to_the_generic_one(det, num_samples, frac):
    g = gen(11)
    rpn, hw, levels, at, grid = _pyramid(det, 448)
    gts = _gts(10, g)
    gts[0] = torch.zeros(0, 4)                      # no positives at all
    gts[1] = rand_boxes(300, 448.0, g, 0.5)         # very many positives: list overflow / dense positives -> generic code
    gts[2] = rand_boxes(3, 448.0, g, 0.02)          # tiny boxes: a handful of positives
    dev_gts = [b.cuda() for b in gts]
    matched, labels, stats, table, off = rpn.anchor_matcher.match_boxes(dev_gts, at, grid=grid, with_stats=True)
    for seed in (1, 2):
        want = det.subsample_labels_(labels.clone(), num_samples, frac, seed)
        got, samples, counts = det.subsample_labels_(labels.clone(), num_samples, frac, seed, stats=stats,
                                                     return_samples=True)
        assert torch.equal(got, want)
        for i in range(len(gts)):
            k = int(counts[i])
            keep = (want[i] != -1).nonzero()[:, 0].cpu()
            assert k == keep.numel() <= num_samples
            lst = samples[i, :k].cpu()
            order = (lst & 0xffffff).argsort()
            assert torch.equal((lst & 0xffffff)[order], keep.to(torch.int32))
            assert torch.equal((lst >> 24)[order].to(torch.int8), want[i].cpu()[keep])


def _heads(n, img, g, scale=0.5):
    obj = [torch.randn(n, 3, img // s, img // s, generator=g) for s in STRIDES]
    dlt = [torch.randn(n, 12, img // s, img // s, generator=g) * scale for s in STRIDES]
    return obj, dlt


@pytest.mark.parametrize("loss_type,beta", [("smooth_l1", 0.0), ("smooth_l1", 0.11), ("giou", 0.0)])
def test_sampled_loss_on_nchw_heads_vs_oracle(det, O, loss_type, beta):
    """Loss and gradients from the conv-layout head == the oracle on the re-laid-out tensors (reference
    rpn.py:270-284 + :187-244), 1e-5."""
    g = gen(21)
    n, img = 4, 224
    rpn = det.RegionProposalNetwork(STRIDES, SIZES, RATIOS, box_reg_loss_type=loss_type, smooth_l1_beta=beta)
    hw = [(img // s, img // s) for s in STRIDES]
    at = torch.cat(rpn.anchor_generator.grid_anchors(hw, torch.device("cuda")), 0)
    grid = rpn.anchor_generator.grid_layout(hw)
    gts = _gts(n, g, float(img))
    obj, dlt = _heads(n, img, g, 0.3)
    asg = rpn.assign(at, [b.cuda() for b in gts], seed=5, grid=grid)
    labels, matched = asg.labels.cpu(), asg.matched.cpu()
    # oracle on the (h w a) layout
    o_c = [o.clone().requires_grad_(True) for o in obj]
    d_c = [d.clone().requires_grad_(True) for d in dlt]
    flat = [O.head_to_hwa(o, d) for o, d in zip(o_c, d_c)]
    lg = torch.cat([f[0] for f in flat], 1)
    dl = torch.cat([f[1] for f in flat], 1)
    mboxes = torch.stack([gts[i][matched[i]] for i in range(n)])
    want = O.rpn_losses(at.cpu(), lg, labels, dl, mboxes, box_reg_loss_type=loss_type, smooth_l1_beta=beta)
    (want["cls_loss"] + 2.0 * want["loc_loss"]).backward()
    # (a) autograd through the sampled kernel
    o_g = [o.cuda().requires_grad_(True) for o in obj]
    d_g = [d.cuda().requires_grad_(True) for d in dlt]
    res = rpn.sampled_losses(at, o_g, d_g, asg)
    (res["cls_loss"] + 2.0 * res["loc_loss"]).backward()
    tol = 2e-5 if loss_type == "giou" else 1e-5
    torch.testing.assert_close(res["cls_loss"].cpu(), want["cls_loss"].detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(res["loc_loss"].cpu(), want["loc_loss"].detach(), rtol=tol, atol=1e-7)
    assert int(res["num_pos_anchors"]) == want["num_pos"] and int(res["num_neg_anchors"]) == want["num_neg"]
    gtol = 1e-4 if loss_type == "giou" else 1e-5
    for a, b in zip(o_g, o_c):
        torch.testing.assert_close(a.grad.cpu(), b.grad, rtol=1e-5, atol=1e-9)
    for a, b in zip(d_g, d_c):
        torch.testing.assert_close(a.grad.cpu(), b.grad, rtol=gtol, atol=1e-9)
    # (b) one launch, forward + backward into caller-owned buffers == the dense fused kernel on flat tensors
    g_obj = [torch.zeros_like(o) for o in o_g]
    g_dlt = [torch.zeros_like(d) for d in d_g]
    res_b = rpn.sampled_losses(at, o_g, d_g, asg, grad_buffers=(g_obj, g_dlt))
    dense = rpn.fused_losses(at, lg.detach().cuda(), dl.detach().cuda(), asg, with_grads=True)
    torch.testing.assert_close(res_b["sums"][:4], dense["sums"][:4], rtol=1e-5, atol=1e-7)
    got_gl = torch.cat([x.permute(0, 2, 3, 1).reshape(n, -1) for x in g_obj], 1)
    got_gd = torch.cat([x.view(n, 3, 4, x.shape[2], x.shape[3]).permute(0, 3, 4, 1, 2).reshape(n, -1, 4) for x in g_dlt], 1)
    assert torch.equal(got_gl, dense["grad_logits"]) and torch.equal(got_gd, dense["grad_deltas"])


def test_sampled_loss_persistent_buffers_clear_the_previous_step(det):
    """Persistent gradient buffers: passing the previous step's sample list resets exactly what that step wrote, so the
    buffers equal freshly zeroed ones after every step -- including anchors sampled in both steps."""
    g = gen(31)
    n, img = 3, 224
    rpn = det.RegionProposalNetwork(STRIDES, SIZES, RATIOS)
    hw = [(img // s, img // s) for s in STRIDES]
    at = torch.cat(rpn.anchor_generator.grid_anchors(hw, torch.device("cuda")), 0)
    grid = rpn.anchor_generator.grid_layout(hw)
    gts = [b.cuda() for b in _gts(n, g, float(img))]
    obj, dlt = _heads(n, img, g)
    obj, dlt = [o.cuda() for o in obj], [d.cuda() for d in dlt]
    p_obj, p_dlt = [torch.zeros_like(o) for o in obj], [torch.zeros_like(d) for d in dlt]
    prev = None
    for step in range(4):
        asg = rpn.assign(at, gts, seed=40 + step // 2, grid=grid)  # seeds repeat: identical samples in steps (0,1), (2,3)
        rpn.sampled_losses(at, obj, dlt, asg, grad_buffers=(p_obj, p_dlt), clear_previous=prev)
        f_obj, f_dlt = [torch.zeros_like(o) for o in obj], [torch.zeros_like(d) for d in dlt]
        rpn.sampled_losses(at, obj, dlt, asg, grad_buffers=(f_obj, f_dlt))
        for a, b in zip(p_obj + p_dlt, f_obj + f_dlt):
            assert torch.equal(a, b), step
        prev = (asg.samples, asg.sample_count)


def test_sampled_loss_flat_layout_and_empty_images(det):
    g = gen(41)
    n, img = 3, 224
    rpn = det.RegionProposalNetwork(STRIDES, SIZES, RATIOS)
    hw = [(img // s, img // s) for s in STRIDES]
    at = torch.cat(rpn.anchor_generator.grid_anchors(hw, torch.device("cuda")), 0)
    grid = rpn.anchor_generator.grid_layout(hw)
    gts = _gts(n, g, float(img))
    gts[1] = torch.zeros(0, 4)
    asg = rpn.assign(at, [b.cuda() for b in gts], seed=9, grid=grid)
    r = at.shape[0]
    logits = torch.randn(n, r, generator=g).cuda()
    deltas = (torch.randn(n, r, 4, generator=g) * 0.4).cuda()
    dense = rpn.fused_losses(at, logits, deltas, asg, with_grads=True)
    gl, gd = torch.zeros_like(logits), torch.zeros_like(deltas)
    sums = rpn._run_sampled(at, logits, deltas, asg, n, None, (gl, gd))
    torch.testing.assert_close(sums[:4], dense["sums"][:4], rtol=1e-5, atol=1e-7)
    assert torch.equal(gl, dense["grad_logits"]) and torch.equal(gd, dense["grad_deltas"])


def test_forward_training_launches_match_the_oracle_losses(det, O):
    """RegionProposalNetwork.forward(training) from NCHW heads: losses == oracle on the labels the step sampled."""
    g = gen(51)
    n, img = 2, 224
    obj, dlt = _heads(n, img, g, 0.3)
    gts = _gts(n, g, float(img))
    insts = []
    for b in gts:
        inst = det.Instances((img, img))
        inst.gt_boxes = det.Boxes(b.cuda())
        insts.append(inst)
    rpn = det.RegionProposalNetwork(STRIDES, SIZES, RATIOS).train()
    o_g = [o.cuda().requires_grad_(True) for o in obj]
    d_g = [d.cuda().requires_grad_(True) for d in dlt]
    rpn._sample_seed = 76  # forward() draws seed 77
    props, losses = rpn.forward([(img, img)] * n, head_outputs=(o_g, d_g), gt_instances=insts)
    (losses["cls_loss"] + losses["loc_loss"]).backward()
    hw = [(img // s, img // s) for s in STRIDES]
    at = torch.cat(rpn.anchor_generator.grid_anchors(hw, torch.device("cuda")), 0)
    asg = rpn.assign(at, [b.cuda() for b in gts], seed=77, grid=rpn.anchor_generator.grid_layout(hw))
    labels, matched = asg.labels.cpu(), asg.matched.cpu()
    o_c = [o.clone().requires_grad_(True) for o in obj]
    d_c = [d.clone().requires_grad_(True) for d in dlt]
    flat = [O.head_to_hwa(o, d) for o, d in zip(o_c, d_c)]
    mboxes = torch.stack([gts[i][matched[i]] for i in range(n)])
    want = O.rpn_losses(at.cpu(), torch.cat([f[0] for f in flat], 1), labels, torch.cat([f[1] for f in flat], 1), mboxes)
    (want["cls_loss"] + want["loc_loss"]).backward()
    torch.testing.assert_close(losses["cls_loss"].detach().cpu(), want["cls_loss"].detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(losses["loc_loss"].detach().cpu(), want["loc_loss"].detach(), rtol=1e-5, atol=1e-7)
    for a, b in zip(o_g + d_g, o_c + d_c):
        torch.testing.assert_close(a.grad.cpu(), b.grad, rtol=1e-5, atol=1e-9)


def _sample_sets(samples, counts, gt_idx=None):
    out = []
    for i in range(samples.shape[0]):
        k = int(counts[i])
        row = samples[i, :k].cpu()
        order = (row & 0xffffff).argsort()
        rec = [(row & 0xffffff)[order], (row >> 24).to(torch.int8)[order]]
        if gt_idx is not None:
            rec.append(gt_idx[i, :k].cpu()[order])
        out.append(rec)
    return out


@pytest.mark.parametrize("num_samples,frac", [(256, 0.5), (64, 0.25)])
def test_sample_list_only_assignment_equals_the_dense_one(det, O, num_samples, frac):
    """det_assign_sampled (nothing dense is computed) picks exactly the anchors det_match_grid +
    det_subsample_labels_grid pick, with the same labels and the same matched gt -- including images without gt, with a
    gt off the image / of zero area (every anchor promoted: dense path), with > 32 gts and with hundreds of positives."""
    g = gen(61)
    rpn, hw, levels, at, grid = _pyramid(det, 448)
    rpn.batch_size_per_image, rpn.positive_fraction = num_samples, frac
    gts = _gts(12, g)
    gts[0] = torch.zeros(0, 4)
    gts[1] = torch.cat([gts[1], torch.tensor([[-500.0, -500.0, -400.0, -450.0]])])   # row maximum 0: everything positive
    gts[2] = torch.cat([gts[2], torch.tensor([[50.0, 60.0, 50.0, 90.0]])])           # zero area
    gts[3] = rand_boxes(70, 448.0, g)                                                # > 32 gts
    gts[4] = rand_boxes(300, 448.0, g, 0.5)                                          # positives overflow the list
    gts[5] = rand_boxes(3, 448.0, g, 0.02)                                           # tiny boxes: low-quality matches only
    gts[6] = torch.cat([gts[6], gts[6][:2]])                                         # duplicated gts
    dev_gts = [b.cuda() for b in gts]
    for seed in (3, 4):
        dense = rpn.assign(at, dev_gts, seed=seed, grid=grid)
        table, off = rpn.anchor_matcher.pack_gt(dev_gts, at.device)
        lazy = rpn.assign_sampled(at, table, off, len(gts), grid, seed=seed)
        assert lazy.labels is None and lazy.matched is None
        assert torch.equal(lazy.sample_count, dense.sample_count)
        want = _sample_sets(dense.samples, dense.sample_count)
        got = _sample_sets(lazy.samples, lazy.sample_count, lazy.sample_gt)
        for i in range(len(gts)):
            assert torch.equal(got[i][0], want[i][0]) and torch.equal(got[i][1], want[i][1]), i
            rows = got[i][0].long()
            pos = got[i][1] == 1
            assert torch.equal(got[i][2][pos].long(), dense.matched[i].cpu()[rows][pos]), i
            # and the dense labels agree with the lists
            lab = dense.labels[i].cpu()
            assert torch.equal(lab[rows], got[i][1]) and int((lab != -1).sum()) == rows.numel()


def test_sample_list_only_assignment_feeds_the_same_losses(det):
    g = gen(62)
    n, img = 4, 224
    rpn = det.RegionProposalNetwork(STRIDES, SIZES, RATIOS)
    hw = [(img // s, img // s) for s in STRIDES]
    at = torch.cat(rpn.anchor_generator.grid_anchors(hw, torch.device("cuda")), 0)
    grid = rpn.anchor_generator.grid_layout(hw)
    gts = [b.cuda() for b in _gts(n, g, float(img))]
    obj, dlt = _heads(n, img, g, 0.3)
    obj, dlt = [o.cuda() for o in obj], [d.cuda() for d in dlt]
    dense = rpn.assign(at, gts, seed=8, grid=grid)
    table, off = rpn.anchor_matcher.pack_gt(gts, at.device)
    lazy = rpn.assign_sampled(at, table, off, n, grid, seed=8)
    ga, gb = ([torch.zeros_like(o) for o in obj], [torch.zeros_like(d) for d in dlt]), \
             ([torch.zeros_like(o) for o in obj], [torch.zeros_like(d) for d in dlt])
    ra = rpn.sampled_losses(at, obj, dlt, dense, grad_buffers=ga)
    rb = rpn.sampled_losses(at, obj, dlt, lazy, grad_buffers=gb)
    torch.testing.assert_close(ra["sums"][:4], rb["sums"][:4], rtol=1e-6, atol=1e-8)
    for x, y in zip(ga[0] + ga[1], gb[0] + gb[1]):
        assert torch.equal(x, y)
    # autograd form on the lazy assignment
    o_g = [o.clone().requires_grad_(True) for o in obj]
    d_g = [d.clone().requires_grad_(True) for d in dlt]
    res = rpn.sampled_losses(at, o_g, d_g, lazy)
    (res["cls_loss"] + res["loc_loss"]).backward()
    for x, y in zip(o_g + d_g, gb[0] + gb[1]):
        assert torch.equal(x.grad, y)


@pytest.mark.parametrize("seed", list(range(int(__import__("os").environ.get("DET_STRESS_SEEDS", "24")))))
def test_window_logic_stress_random_pyramids(det, seed):
    """The closed-form windows of the grid kernels (reach / plateau / overlap windows, one position of margin) against the
    brute-force generic matcher on random pyramids: strides, anchor sizes, aspect ratios, anchor offset 0 or 0.5,
    non-square ragged feature maps, and gt boxes from sub-pixel to larger than the image, partly outside it.  The dense
    grid matcher must equal det_match_anchors bit for bit; the sample-list assignment must equal the dense one."""
    g = gen(1000 + seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    nlev = ri(1, 4)
    s0 = [2, 4, 8][ri(0, 2)]
    strides = [s0 * (2 ** l) for l in range(nlev)]
    base = float(torch.rand(1, generator=g)) * 24 + 8
    sizes = [[base * (2 ** l)] for l in range(nlev)]
    ratios = [[[0.5, 1.0, 2.0], [1.0], [0.33, 1.0, 3.0]][ri(0, 2)]]
    offset = [0.0, 0.5][ri(0, 1)]
    H, W = ri(40, 260), ri(40, 260)
    rpn = det.RegionProposalNetwork(strides, sizes, ratios, anchor_offset=offset)
    hw = [(-(-H // s), -(-W // s)) for s in strides]
    at = torch.cat(rpn.anchor_generator.grid_anchors(hw, torch.device("cuda")), 0)
    grid = rpn.anchor_generator.grid_layout(hw)
    n = 6
    gts = []
    for i in range(n):
        k = ri(0, 12)
        scale = [2.0, 20.0, 80.0, 400.0][ri(0, 3)]
        xy = torch.rand(k, 2, generator=g) * torch.tensor([W * 1.2, H * 1.2]) - torch.tensor([W * 0.1, H * 0.1])
        wh = torch.rand(k, 2, generator=g) * scale + 0.05
        gts.append(torch.cat([xy, xy + wh], 1).cuda())
    dense = rpn.assign(at, gts, sample=False, grid=grid)
    generic = rpn.assign(at, gts, sample=False)
    assert torch.equal(dense.labels, generic.labels) and torch.equal(dense.matched, generic.matched)
    rpn.batch_size_per_image = 64
    d2 = rpn.assign(at, gts, seed=5, grid=grid)
    table, off = rpn.anchor_matcher.pack_gt(gts, at.device)
    lazy = rpn.assign_sampled(at, table, off, n, grid, seed=5)
    assert torch.equal(lazy.sample_count, d2.sample_count)
    want = _sample_sets(d2.samples, d2.sample_count)
    got = _sample_sets(lazy.samples, lazy.sample_count, lazy.sample_gt)
    for i in range(n):
        assert torch.equal(got[i][0], want[i][0]) and torch.equal(got[i][1], want[i][1]), i
        rows, pos = got[i][0].long(), got[i][1] == 1
        assert torch.equal(got[i][2][pos].long(), d2.matched[i].cpu()[rows][pos]), i
