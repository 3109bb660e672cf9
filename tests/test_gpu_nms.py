"""GPU parity: batched NMS kernels vs the CPU oracle -- kept indices and counts bit-exact."""
import pytest
import torch

from tests.util import gen, rand_boxes, distinct_scores

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def det():
    import det_b200
    return det_b200


@pytest.fixture(scope="module")
def O():
    from oracle import ref_torch
    return ref_torch


def _case(n, ncat, seed, frame=640.0, quant=False, ties=False):
    g = gen(seed)
    b = rand_boxes(n, frame, g)
    if quant:
        b = b.round()
    s = torch.rand(n, generator=g)
    if ties:
        s = (s * 16).round() / 16
    c = torch.randint(0, ncat, (n,), generator=g)
    return b, s, c


@pytest.mark.parametrize("n,ncat", [(1, 1), (2, 1), (31, 3), (100, 1), (257, 5), (900, 80), (1000, 20), (1001, 20),
                                     (1024, 7), (1960, 20), (2048, 80), (3000, 80), (4096, 1), (4097, 3),
                                     (6000, 5), (25200, 80), (25200, 1)])
@pytest.mark.parametrize("thr", [0.5, 0.7])
def test_batched_nms_matches_oracle(det, O, n, ncat, thr):
    b, s, c = _case(n, ncat, seed=n * 7 + ncat, quant=(n % 2 == 0), ties=(n % 3 == 0))
    want = O.batched_nms(b, s, c, thr)
    got = det.batched_nms(b.cuda(), s.cuda(), c.cuda(), thr).cpu()
    assert got.dtype == torch.int64
    assert got.numel() == want.numel()
    assert torch.equal(got, want)


@pytest.mark.parametrize("n", [50, 700, 5000])
def test_nms_single_category_vs_torchvision_semantics(det, O, n):
    b, s, _ = _case(n, 1, seed=n, frame=200.0)
    assert torch.equal(det.nms(b.cuda(), s.cuda(), 0.3).cpu(), O.nms(b, s, 0.3))


def test_threshold_is_strict_and_double(det, O):
    # IoU exactly 0.5 is kept (strict >); IoU == float32(0.7) is kept against the double 0.7
    b = torch.tensor([[0, 0, 2, 1], [0, 0, 1, 1.0]])
    s = torch.tensor([1.0, 0.5])
    assert det.nms(b.cuda(), s.cuda(), 0.5).tolist() == [0, 1]
    b = torch.tensor([[0, 0, 10, 1], [0, 0, 7, 1.0]])
    assert det.nms(b.cuda(), s.cuda(), 0.7).tolist() == O.nms(b, s, 0.7).tolist() == [0, 1]
    assert det.nms(b.cuda(), s.cuda(), 0.6999).tolist() == O.nms(b, s, 0.6999).tolist() == [0]


def test_nan_scores_and_degenerate_boxes(det, O):
    g = gen(3)
    b = rand_boxes(300, 100.0, g)
    s = torch.rand(300, generator=g)
    s[::17] = float("nan")
    b[5] = torch.tensor([1.0, 1.0, 1.0, 1.0])
    b[6] = torch.tensor([1.0, 1.0, 1.0, 1.0])
    c = torch.randint(0, 3, (300,), generator=g)
    for mode_n in (300,):
        assert torch.equal(det.batched_nms(b.cuda(), s.cuda(), c.cuda(), 0.5).cpu(), O.batched_nms(b, s, c, 0.5))


def test_offset_trick_with_negative_coordinates(det, O):
    # coordinates below -1 can make boxes of different categories overlap after the offset: general path
    g = gen(9)
    b = rand_boxes(400, 300.0, g) - 150.0
    s = torch.rand(400, generator=g)
    c = torch.randint(0, 4, (400,), generator=g)
    assert torch.equal(det.batched_nms(b.cuda(), s.cuda(), c.cuda(), 0.3).cpu(), O.batched_nms(b, s, c, 0.3))


def test_images_batch_with_counts(det, O):
    g = gen(21)
    n_img, m = 9, 1500
    boxes = torch.stack([rand_boxes(m, 400.0, g) for _ in range(n_img)])
    scores = torch.rand(n_img, m, generator=g)
    cats = torch.randint(0, 20, (n_img, m), generator=g)
    counts = torch.tensor([1500, 0, 1, 999, 1000, 1001, 1499, 64, 65], dtype=torch.int32)
    keep, kc = det.nms_images(boxes.cuda(), scores.cuda(), cats.cuda(), counts.cuda(), 0.5)
    keep, kc = keep.cpu(), kc.cpu()
    for i in range(n_img):
        k = int(counts[i])
        want = O.batched_nms(boxes[i, :k], scores[i, :k], cats[i, :k], 0.5)
        assert int(kc[i]) == want.numel()
        assert torch.equal(keep[i, :want.numel()], want)


def test_images_batch_large_path_max_out(det, O):
    g = gen(22)
    n_img, m = 3, 7000
    boxes = torch.stack([rand_boxes(m, 500.0, g) for _ in range(n_img)])
    scores = torch.rand(n_img, m, generator=g)
    cats = torch.randint(0, 5, (n_img, m), generator=g)
    counts = torch.tensor([7000, 4500, 10], dtype=torch.int32)
    keep, kc = det.nms_images(boxes.cuda(), scores.cuda(), cats.cuda(), counts.cuda(), 0.6, max_out=300)
    keep, kc = keep.cpu(), kc.cpu()
    for i in range(n_img):
        k = int(counts[i])
        want = O.batched_nms(boxes[i, :k], scores[i, :k], cats[i, :k], 0.6)[:300]
        assert int(kc[i]) == want.numel()
        assert torch.equal(keep[i, :want.numel()], want)


def test_bad_category_raises(det):
    b = torch.tensor([[0, 0, 1, 1.0]]).cuda()
    with pytest.raises(ValueError):
        det.batched_nms(b, torch.tensor([1.0]).cuda(), torch.tensor([-3]).cuda(), 0.5)


# ---- long segments: CTA path with shared-memory staging (<= 4096) and the cooperative grid-wide path (> 4096) -------
@pytest.mark.parametrize("m,ncat,thr,max_out", [
    (4096, 1, 0.5, None),      # exactly the CTA-path limit
    (4097, 1, 0.5, None),      # first huge segment
    (9000, 1, 0.4, None),      # several 512-box leader steps, heavy suppression
    (9000, 1, 0.7, 700),       # early stop at max_out in the middle of a leader chunk
    (12000, 2, 0.5, None),     # two huge segments per image advancing in lock step
    (6000, 3, 0.5, None),      # mid segments only
])
def test_long_segments_match_oracle(det, O, m, ncat, thr, max_out):
    g = gen(31 + m % 7)
    n_img = 3
    boxes = torch.stack([rand_boxes(m, 700.0, g, 0.25) for _ in range(n_img)])
    scores = torch.stack([distinct_scores(m, g) for _ in range(n_img)])
    cats = torch.randint(0, ncat, (n_img, m), generator=g)
    counts = torch.tensor([m, m - 1234, 37], dtype=torch.int32)
    keep, kc = det.nms_images(boxes.cuda(), scores.cuda(), cats.cuda(), counts.cuda(), thr, max_out=max_out, mode=1)
    keep, kc = keep.cpu(), kc.cpu()
    for i in range(n_img):
        k = int(counts[i])
        want = O.batched_nms(boxes[i, :k], scores[i, :k], cats[i, :k], thr)
        if max_out is not None:
            want = want[:max_out]
        assert int(kc[i]) == want.numel(), (i, int(kc[i]), want.numel())
        assert torch.equal(keep[i, :want.numel()], want), i


def test_huge_segment_with_nan_and_tied_scores(det, O):
    """NaN coordinates switch the image to the NaN-safe predicate; exactly tied scores must come out in index order."""
    g = gen(41)
    m = 5200
    boxes = rand_boxes(m, 600.0, g, 0.3)
    boxes[17, 2] = float("nan")
    boxes[4000, 0] = float("nan")
    scores = torch.randint(0, 400, (m,), generator=g).float() / 400.0  # many exact ties
    keep, kc = det.nms_images(boxes[None].cuda(), scores[None].cuda(), None, None, 0.5, mode=1)
    want = O.nms(boxes, scores, 0.5)
    assert int(kc[0]) == want.numel() and torch.equal(keep[0, :want.numel()].cpu(), want)


# ---- BASELINE configs[4] sizes against the oracle itself (kept indices + counts bit-exact).  40 500 boxes crosses the
#      reference's own >= 40000 branch (python/src/utils.py:108-119: per-id loop + argsort instead of torchvision's
#      batched_nms); torch.rand fp32 scores tie at these sizes, ties resolve by lower index.
@pytest.mark.parametrize("m,ncat,frame,seed", [
    (40500, 30, 1024.0, 11),    # the reference's branch point
    (50000, 80, 1024.0, 12),
    (100000, 80, 1024.0, 13),
    (50000, 1, 1024.0, 14),     # single-category worst case: one 50k-box segment
    (100000, 3, 1024.0, 15),    # three ~33k-box segments (cooperative huge-segment sweep)
])
def test_batched_nms_matches_oracle_at_cfg5_sizes(det, O, m, ncat, frame, seed):
    g = gen(seed)
    b = rand_boxes(m, frame, g, 0.2)
    s = torch.rand(m, generator=g)
    c = torch.randint(0, ncat, (m,), generator=g)
    want = O.batched_nms(b, s, c, 0.5)
    got = det.batched_nms(b.cuda(), s.cuda(), c.cuda(), 0.5).cpu()
    assert got.numel() == want.numel()
    assert torch.equal(got, want)


def test_nms_images_cfg5_batch_of_8_at_50k(det, O):
    """configs[4] as it is sharded: 8 images per GPU in one call, ragged counts around 50 000 boxes, 80 categories."""
    n, m = 8, 50000
    g = gen(77)
    counts = [m, m - 1, 40000, 39999, 45000, 50000, 12345, 1000]
    bs = rand_boxes(n * m, 1024.0, g, 0.2).view(n, m, 4)
    ss = torch.rand(n, m, generator=g)
    cs = torch.randint(0, 80, (n, m), generator=g)
    keep, cnt = det.nms_images(bs.cuda(), ss.cuda(), cs.cuda(), torch.tensor(counts, dtype=torch.int32).cuda(), 0.5)
    for i in range(n):
        k = counts[i]
        want = O.batched_nms(bs[i, :k], ss[i, :k], cs[i, :k], 0.5)
        assert int(cnt[i]) == want.numel()
        assert torch.equal(keep[i, :want.numel()].cpu(), want)


def test_nms_properties_at_100k_boxes(det, O):
    """BASELINE configs[4] size: size-independent properties of greedy NMS (the oracle-equality cases are above).
    (1) kept indices are unique, valid and in descending-score order; (2) idempotence: NMS of the kept set keeps
    everything, in the same order; (3) no kept box is suppressed by an earlier kept box (sampled pairs);
    (4) every dropped box of a sample is suppressed by some kept box with a higher score."""
    g = gen(51)
    m = 100000
    boxes = rand_boxes(m, 1024.0, g, 0.2)
    scores = distinct_scores(m, g)
    cats = torch.randint(0, 3, (m,), generator=g)
    b, s, c = boxes.cuda(), scores.cuda(), cats.cuda()
    keep, kc = det.nms_images(b[None], s[None], c[None], None, 0.5, mode=1)
    k = int(kc[0])
    kept = keep[0, :k]
    assert 0 < k < m and int(kept.min()) >= 0 and int(kept.max()) < m and kept.unique().numel() == k
    ks = s[kept]
    assert bool((ks[1:] < ks[:-1]).all())
    keep2, kc2 = det.nms_images(b[kept][None], ks[None], c[kept][None], None, 0.5, mode=1)
    assert int(kc2[0]) == k and torch.equal(keep2[0, :k], torch.arange(k, device=keep2.device))
    # (3) on a random sample of kept boxes vs all kept boxes of the same category
    kb, kcat = b[kept], c[kept]
    sel = torch.randperm(k, generator=g)[:300].cuda()
    iou = det.pairwise_iou(det.Boxes(kb[sel]), det.Boxes(kb))
    same = kcat[sel][:, None] == kcat[None, :]
    iou = torch.where(same, iou, torch.zeros_like(iou))
    iou[torch.arange(sel.numel(), device=iou.device), sel] = 0
    assert float(iou.max()) <= 0.5
    # (4) dropped boxes: each has a same-category kept box with higher score and IoU > 0.5
    dropped = torch.ones(m, dtype=torch.bool, device=b.device)
    dropped[kept] = False
    didx = dropped.nonzero()[:, 0]
    didx = didx[torch.randperm(didx.numel(), generator=g)[:300].cuda()]
    iou = det.pairwise_iou(det.Boxes(b[didx]), det.Boxes(kb))
    ok = (iou > 0.5) & (c[didx][:, None] == kcat[None, :]) & (ks[None, :] > s[didx][:, None])
    assert bool(ok.any(dim=1).all())


# ---- offset trick with coordinates below -1: categories are swept independently only if no pair of boxes of different
#      categories intersects after the shift (nms_small.cuh phase 0) ------------------------------------------------------
@pytest.mark.parametrize("n,ncat,seed", [(300, 80, 0), (900, 20, 1), (1000, 80, 2), (64, 2, 3), (700, 3, 4)])
def test_offset_trick_with_negative_coordinates(det, O, n, ncat, seed):
    g = gen(100 + seed)
    b = rand_boxes(n, 640.0, g, wh_frac=0.4) - 120.0  # a good share of the boxes starts left of / above the frame
    s = torch.rand(n, generator=g)
    c = torch.randint(0, ncat, (n,), generator=g)
    for thr in (0.5, 0.05):
        want = O.batched_nms(b, s, c, thr)
        got = det.batched_nms(b.cuda(), s.cuda(), c.cuda(), thr).cpu()
        assert torch.equal(got, want)


def test_offset_trick_cross_category_suppression_is_reproduced(det, O):
    """torchvision's trick lets a top-left box of category c+1 overlap a bottom-right box of category c: the shifted
    boxes really intersect and the lower-scored one is suppressed ACROSS categories.  The library must do the same."""
    g = gen(7)
    b = rand_boxes(200, 400.0, g)
    s = torch.rand(200, generator=g) * 0.5
    c = torch.randint(0, 4, (200,), generator=g)
    mx = 500.0
    b[0] = torch.tensor([mx - 60, mx - 60, mx, mx]); c[0] = 0; s[0] = 0.9      # holds the global maximum coordinate
    b[1] = torch.tensor([-55.0, -55.0, 8.0, 8.0]); c[1] = 1; s[1] = 0.8        # shifted by span = mx + 1 it meets box 0
    want = O.batched_nms(b, s, c, 0.05)
    assert 1 not in want.tolist() and 0 in want.tolist()  # suppressed across categories in the reference semantics
    got = det.batched_nms(b.cuda(), s.cuda(), c.cuda(), 0.05).cpu()
    assert torch.equal(got, want)
    # the same boxes per category (forced branch) keep box 1
    kk, kc = det.nms_images(b[None].cuda(), s[None].cuda(), c[None].cuda(), None, 0.05, None, 1)
    assert 1 in kk[0, :int(kc)].tolist()


def test_offset_trick_many_corner_boxes_take_the_exact_slow_path(det, O):
    g = gen(8)
    n = 600
    b = rand_boxes(n, 300.0, g) - 200.0  # most boxes have x1, y1 < -1: more than kCrossMax candidates
    s = torch.rand(n, generator=g)
    c = torch.randint(0, 10, (n,), generator=g)
    assert torch.equal(det.batched_nms(b.cuda(), s.cuda(), c.cuda(), 0.4).cpu(), O.batched_nms(b, s, c, 0.4))


# ---- top-k tier of the large path: max_out << boxes (nms_large.cuh large_topk_select_kernel + nms_list.cuh) ------------
@pytest.mark.parametrize("m,ncat,max_out,thr,ties", [
    (5000, 80, 100, 0.5, False), (25200, 80, 1000, 0.5, False), (25200, 1, 300, 0.5, False),
    (8000, 20, 200, 0.5, True),      # quantised scores: ties inside segments, across them and at the radix cut
    (6000, 1, 500, 0.05, False),     # heavy suppression: the tier falls short, the full path takes over
    (5000, 3, 2400, 0.5, False),     # the largest tier (want = 3032)
])
def test_large_path_topk_tier(det, O, m, ncat, max_out, thr, ties):
    n = 3
    counts = [m, m // 2 + 7, max_out + 5]  # the last image is too small for a tier
    bs, ss, cs = [], [], []
    for i in range(n):
        b, s, c = _case(m, ncat, seed=1000 + m + i, ties=ties)
        bs.append(b); ss.append(s); cs.append(c)
    keep, cnt = det.nms_images(torch.stack(bs).cuda(), torch.stack(ss).cuda(), torch.stack(cs).cuda(),
                               torch.tensor(counts, dtype=torch.int32).cuda(), thr, max_out)
    for i in range(n):
        k = counts[i]
        want = O.batched_nms(bs[i][:k], ss[i][:k], cs[i][:k], thr)[:max_out]
        assert int(cnt[i]) == want.numel()
        assert torch.equal(keep[i, :want.numel()].cpu(), want)


@pytest.mark.parametrize("mode", [1, 2])
def test_large_path_topk_tier_forced_branches(det, mode):
    """Forced per-category / offset-trick branches: top-k result == the first max_out of the library's full result."""
    b, s, c = _case(7000, 12, seed=5)
    b = b - 100.0  # coordinates below -1: the trick's cross-category handling with the full image's span
    args = (b[None].cuda(), s[None].cuda(), c[None].cuda(), None, 0.5)
    full, fc = det.nms_images(*args, None, mode)
    top, tc = det.nms_images(*args, 150, mode)
    assert int(tc) == 150 and int(fc) >= 150 and torch.equal(top[0, :150], full[0, :150])


def test_large_path_topk_tier_bad_category_anywhere_is_reported(det):
    b, s, c = _case(6000, 5, seed=6)
    s[4000] = -1.0   # far below the tier's cut
    c[4000] = 40000  # outside the key range
    keep, cnt = det.nms_images(b[None].cuda(), s[None].cuda(), c[None].cuda(), None, 0.5, 100)
    assert int(cnt) == -1


@pytest.mark.parametrize("seed", list(range(int(__import__("os").environ.get("DET_STRESS_SEEDS", "16")))))
def test_nms_images_stress_random_batches(det, O, seed):
    """Random batches through det_nms_batched: image counts around every path boundary (warp / CTA / large path, the
    1000-box branch rule), 1..200 categories with skewed sizes, frames from crowded to empty, negative coordinates, quantised
    coordinates and scores (ties), an occasional NaN / degenerate box, random max_out: kept indices and counts bit-exact."""
    g = gen(9000 + seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    rf = lambda lo, hi: float(torch.rand(1, generator=g)) * (hi - lo) + lo
    n_img = ri(1, 5)
    m = [40, 300, 1100, 4200, 6000][ri(0, 4)]
    ncat = [1, 2, 7, 80, 200][ri(0, 4)]
    frame = [60.0, 300.0, 2000.0][ri(0, 2)]
    boxes = torch.stack([rand_boxes(m, frame, g, rf(0.05, 0.6)) for _ in range(n_img)]) - (frame * 0.3 if ri(0, 2) == 0 else 0.0)
    if ri(0, 1):
        boxes = boxes.round()
    scores = torch.rand(n_img, m, generator=g)
    if ri(0, 2) == 0:
        scores = (scores * 32).round() / 32
    cats = (torch.rand(n_img, m, generator=g) ** [1.0, 3.0][ri(0, 1)] * ncat).long().clamp(max=ncat - 1)  # skewed sizes
    if ri(0, 3) == 0:
        boxes[0, ri(0, m - 1), ri(0, 3)] = float("nan")
    if ri(0, 3) == 0:
        j = ri(0, m - 1)
        boxes[0, j, 2:] = boxes[0, j, :2]  # zero-area box
    counts = torch.tensor([[m, ri(0, m), min(m, 1000), min(m, 1001), ri(0, m)][i] for i in range(n_img)], dtype=torch.int32)
    thr = rf(0.1, 0.9)
    max_out = [None, 1, 100, 1000][ri(0, 3)]
    keep, kc = det.nms_images(boxes.cuda(), scores.cuda(), cats.cuda(), counts.cuda(), thr, max_out)
    keep, kc = keep.cpu(), kc.cpu()
    for i in range(n_img):
        k = int(counts[i])
        want = O.batched_nms(boxes[i, :k], scores[i, :k], cats[i, :k], thr)
        if max_out is not None:
            want = want[:max_out]
        assert int(kc[i]) == want.numel(), (i, int(kc[i]), want.numel())
        assert torch.equal(keep[i, :want.numel()], want), i


# ---- CTA-class segments through the spatial index (large_bin_segments_kernel; IoU threshold >= 0.5) ---------------------
def _spatial_case(kind, m, ncat, g):
    if kind == "sparse":          # small boxes in a wide frame: nearly everything survives
        b = rand_boxes(m, 4000.0, g, 0.015)
    elif kind == "clustered":     # near-duplicates around a few centres: many suppressors per box, few survivors
        k = 40
        ctr = torch.rand(k, 2, generator=g) * 900 + 50
        which = torch.randint(0, k, (m,), generator=g)
        c = ctr[which] + torch.randn(m, 2, generator=g) * 4.0
        wh = 60 + torch.randn(m, 2, generator=g) * 5.0
        b = torch.cat([c - wh / 2, c + wh / 2], 1)
    elif kind == "chain":         # a staircase: every box is suppressed by its predecessor only -> a dependency chain of m
        i = torch.arange(m, dtype=torch.float32)
        b = torch.stack([i * 15.0, i * 0.0, i * 15.0 + 100.0, i * 0.0 + 50.0], 1)
    elif kind == "mixed":         # boxes of every size, some covering the whole frame
        b = rand_boxes(m, 1000.0, g, 0.1)
        big = torch.randint(0, m, (m // 50,), generator=g)
        b[big] = torch.tensor([0.0, 0.0, 1000.0, 1000.0]) + torch.randn(big.numel(), 4, generator=g) * 20
    elif kind == "quantised":
        b = rand_boxes(m, 300.0, g, 0.2).round()
    elif kind == "tiny":          # coordinates far below the range where the index is trusted: the sweep takes over
        b = rand_boxes(m, 640.0, g, 0.2) * 1e-24
    elif kind == "giant":
        b = rand_boxes(m, 640.0, g, 0.2) * 1e16
    elif kind == "degenerate":    # zero-width, inverted and far-away boxes among ordinary ones
        b = rand_boxes(m, 500.0, g, 0.2)
        j = torch.randperm(m, generator=g)
        b[j[:50], 2] = b[j[:50], 0]
        b[j[50:100], 2] = b[j[50:100], 0] - 5.0
        b[j[100:110]] += 1e6
    else:
        raise AssertionError(kind)
    return b


@pytest.mark.parametrize("kind,m,ncat,thr,max_out,ordered", [
    ("sparse", 6000, 3, 0.7, None, False),
    ("sparse", 6000, 2, 0.5, 1500, False),
    ("clustered", 6000, 4, 0.5, None, False),
    ("clustered", 5000, 2, 0.7, None, False),
    ("clustered", 5000, 2, 0.7, 300, False),
    ("chain", 4500, 9, 0.7, None, True),
    ("chain", 4500, 9, 0.7, None, False),
    ("mixed", 6000, 3, 0.5, None, False),
    ("mixed", 8000, 2, 0.6, None, False),
    ("quantised", 6000, 5, 0.5, None, False),
    ("quantised", 6000, 20, 0.5, None, False),
    ("tiny", 5000, 3, 0.5, None, False),
    ("giant", 5000, 3, 0.5, None, False),
    ("degenerate", 6000, 3, 0.5, None, False),
    ("sparse", 6000, 3, 0.4999, None, False),   # below 1/2: the index must not be used
])
def test_spatial_index_segments_match_oracle(det, O, kind, m, ncat, thr, max_out, ordered):
    g = gen(77 + m % 13 + ncat)
    n_img = 2
    boxes = torch.stack([_spatial_case(kind, m, ncat, g) for _ in range(n_img)])
    if ordered:   # scores fall along the chain: kept, dead, kept, ... decided one box per round
        scores = torch.stack([torch.linspace(1.0, 0.01, m) for _ in range(n_img)])
        cats = (torch.arange(m) * ncat // m)[None].repeat(n_img, 1)
    else:
        scores = torch.stack([distinct_scores(m, g) for _ in range(n_img)])
        cats = torch.randint(0, ncat, (n_img, m), generator=g)
    counts = torch.tensor([m, m - 777], dtype=torch.int32)
    keep, kc = det.nms_images(boxes.cuda(), scores.cuda(), cats.cuda(), counts.cuda(), thr, max_out=max_out, mode=1)
    keep, kc = keep.cpu(), kc.cpu()
    for i in range(n_img):
        k = int(counts[i])
        want = O.batched_nms(boxes[i, :k], scores[i, :k], cats[i, :k], thr)
        if max_out is not None:
            want = want[:max_out]
        assert int(kc[i]) == want.numel(), (i, int(kc[i]), want.numel())
        assert torch.equal(keep[i, :want.numel()], want), i
