"""GPU parity: dense anchor head as a detector runs it -- decode -> score threshold -> per-class NMS -> top max_det
(det_dense_detect, csrc/dense_decode.cu) vs the oracle's dense_select_nms (own spec over the reference's batched_nms,
python/src/utils.py:96-119).  The decode arithmetic itself is pinned by tests/test_gpu_yolo.py (rtol 1e-5); here the
selection + NMS must be BIT-EXACT on the decoded values: rows, counts, boxes, scores and classes."""
import pytest
import torch

from tests.util import gen

pytestmark = pytest.mark.gpu

STRIDES = [8, 16, 32]
WH = [[[10, 13], [16, 30], [33, 23]], [[30, 61], [62, 45], [59, 119]], [[116, 90], [156, 198], [373, 326]]]


@pytest.fixture(scope="module")
def det():
    import det_b200
    return det_b200


@pytest.fixture(scope="module")
def O():
    from oracle import ref_torch
    return ref_torch


def make_heads(n, image, C, seed, obj_bias, strides=STRIDES):
    g = gen(seed)
    heads = [torch.randn(n, 3 * (5 + C), image // s, image // s, generator=g) for s in strides]
    for h in heads:
        h.view(n, 3, 5 + C, h.shape[2], h.shape[3])[:, :, 4] += obj_bias
    return heads


def check_against_oracle(dh, O, heads_gpu, r, thr, iou, max_det):
    boxes, scores, classes = [t.cpu() for t in dh.decode(heads_gpu)]
    idx, gb, gs, gc, cnt = [r[k].cpu() for k in ("idx", "boxes", "scores", "classes", "count")]
    for i in range(boxes.shape[0]):
        wi, wb, ws, wc = O.dense_select_nms(boxes[i], scores[i], classes[i], thr, iou, max_det)
        k = wi.numel()
        assert int(cnt[i]) == k, (i, int(cnt[i]), k)
        assert torch.equal(idx[i, :k], wi)
        bits = lambda t: t.contiguous().view(torch.int32)  # NaN boxes compare by bit pattern
        assert torch.equal(bits(gb[i, :k]), bits(wb)) and torch.equal(bits(gs[i, :k]), bits(ws)) and torch.equal(gc[i, :k], wc)


@pytest.mark.parametrize("gate", [True, False])
@pytest.mark.parametrize("image,n,bias,thr,cap,max_det", [
    (256, 4, -4.0, 0.1, 1024, 300),    # ~2 % pass: <= 1000 candidates -> coordinate-offset branch
    (256, 3, -1.0, 0.25, 2048, 300),   # > 1000 candidates -> per-category branch
    (256, 2, 0.0, 0.3, 4096, 1000),
    (128, 5, -2.0, 0.05, 1024, 7),     # max_det cuts the list
    (64, 2, -4.0, 0.999, 1024, 300),   # (almost) nothing passes
])
def test_dense_detect_parity(det, O, gate, image, n, bias, thr, cap, max_det):
    C = 80
    dh = det.DenseAnchorHead(STRIDES, WH, C)
    heads = [h.cuda() for h in make_heads(n, image, C, 11 + image + n, bias)]
    r = dh.detect_thresholded(heads, thr, 0.5, max_det=max_det, cand_cap=cap, gate=gate, check=False)
    assert int(r["overflow"].item()) == 0
    check_against_oracle(dh, O, heads, r, thr, 0.5, max_det)


@pytest.mark.parametrize("bias,thr,iou,max_det,cap", [
    (0.0, 0.3, 0.5, 100, 4096),    # tier cut succeeds: the best ~160 candidates yield 100 survivors (per-category branch)
    (0.0, 0.3, 0.01, 100, 4096),   # heavy suppression: the tier falls short and everything is swept
    (-2.5, 0.2, 0.5, 100, 1024),   # <= 1000 candidates: offset-trick branch with the full set's span
    (-2.5, 0.2, 0.02, 100, 1024),  # the same, tier falls short
    (1.0, 0.3, 0.3, 50, 4096),
    (0.0, 0.3, 0.5, 1, 4096),
])
def test_dense_detect_tier_cut(det, O, bias, thr, iou, max_det, cap):
    """Only max_det detections are wanted: the kernel first sweeps the ~1.25 * max_det best candidates (exact)."""
    C = 80
    dh = det.DenseAnchorHead(STRIDES, WH, C)
    heads = [h.cuda() for h in make_heads(4, 256, C, 77, bias)]
    r = dh.detect_thresholded(heads, thr, iou, max_det=max_det, cand_cap=cap, check=False)
    assert int(r["overflow"].item()) == 0
    n_cand = (dh.decode(heads)[1] > thr).sum(1)
    assert int(n_cand.min()) >= 3 * (max_det + max_det // 4 + 32) // 2 + 1, n_cand.tolist()  # the tier path runs
    check_against_oracle(dh, O, heads, r, thr, iou, max_det)


def test_dense_detect_tier_cut_with_quantised_scores(det, O):
    """Scores on a coarse grid: many ties inside segments, across segments and at the tier's radix cut."""
    C = 8
    dh = det.DenseAnchorHead(STRIDES, WH, C)
    heads = make_heads(3, 256, C, 31, 0.0)
    for h in heads:
        v = h.view(3, 3, 5 + C, h.shape[2], h.shape[3])
        v[:, :, 4:] = (v[:, :, 4:] * 2).round() / 2  # objectness and class logits in steps of 0.5
    hg = [h.cuda() for h in heads]
    for max_det in (20, 150):
        r = dh.detect_thresholded(hg, 0.3, 0.5, max_det=max_det, cand_cap=4096, check=False)
        assert int(r["overflow"].item()) == 0
        check_against_oracle(dh, O, hg, r, 0.3, 0.5, max_det)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_dense_detect_modes_and_few_classes(det, O, mode):
    C = 3
    dh = det.DenseAnchorHead(STRIDES, WH, C)
    heads = [h.cuda() for h in make_heads(3, 128, C, 5, -1.5)]
    r = dh.detect_thresholded(heads, 0.2, 0.45, max_det=200, cand_cap=2048, mode=mode, check=False)
    assert int(r["overflow"].item()) == 0
    boxes, scores, classes = [t.cpu() for t in dh.decode(heads)]
    for i in range(3):
        cand = torch.nonzero(scores[i] > 0.2, as_tuple=True)[0]
        if mode == 0:
            keep = O.batched_nms(boxes[i][cand], scores[i][cand], classes[i][cand], 0.45)[:200]
        else:  # forced branches: compare with the library's own generic NMS in the same mode
            kk, kc = det.nms_images(boxes[i][cand][None].cuda(), scores[i][cand][None].cuda(),
                                   classes[i][cand][None].cuda(), None, 0.45, 200, mode)
            keep = kk[0, :int(kc)].cpu()
        k = keep.numel()
        assert int(r["count"][i]) == k and torch.equal(r["idx"][i, :k].cpu(), cand[keep])


def test_dense_detect_ties_resolve_by_row(det, O):
    """Identical logits everywhere in a level: every score ties, torch.nonzero order (row index) must decide."""
    C = 4
    dh = det.DenseAnchorHead(STRIDES, WH, C)
    heads = make_heads(2, 128, C, 9, -3.0)
    heads[1][:] = 0.0
    heads[1].view(2, 3, 5 + C, 8, 8)[:, :, 4] = 1.0     # the whole 8x8 level passes with one score
    heads[1].view(2, 3, 5 + C, 8, 8)[:, :, 5 + 2] = 0.5  # class 2 everywhere
    hg = [h.cuda() for h in heads]
    for gate in (True, False):
        r = dh.detect_thresholded(hg, 0.3, 0.5, max_det=300, cand_cap=1024, gate=gate, check=False)
        check_against_oracle(dh, O, hg, r, 0.3, 0.5, 300)
    assert int(r["count"].min()) > 10


def test_dense_detect_nan_and_inf_logits(det, O):
    C = 6
    dh = det.DenseAnchorHead(STRIDES, WH, C)
    heads = make_heads(3, 128, C, 21, -1.0)
    v = heads[0].view(3, 3, 5 + C, 16, 16)
    v[0, 1, 4, 3, 5] = float("nan")    # NaN objectness: NaN score, never a candidate
    v[0, 2, 4, 7, 7] = float("inf")    # objectness 1
    v[1, 0, 5 + 3, 2, 2] = float("nan")  # NaN class logit wins the arg-max, NaN score
    v[1, 1, 2, 4, 4] = float("inf")    # clamped width
    v[2, 0, 4, 9, 9] = 8.0
    v[2, 0, 0, 9, 9] = float("nan")    # candidate with a NaN box
    hg = [h.cuda() for h in heads]
    for gate in (True, False):
        r = dh.detect_thresholded(hg, 0.15, 0.5, max_det=300, cand_cap=2048, gate=gate, check=False)
        assert int(r["overflow"].item()) == 0
        check_against_oracle(dh, O, hg, r, 0.15, 0.5, 300)


def test_dense_detect_overflow_is_reported_then_redone_exactly(det, O):
    C = 80
    dh = det.DenseAnchorHead(STRIDES, WH, C)
    heads = [h.cuda() for h in make_heads(3, 256, C, 2, 0.0)]
    for h in heads:  # image 1 stays small
        h[1].view(3, 5 + C, h.shape[2], h.shape[3])[:, 4] -= 8.0
    r = dh.detect_thresholded(heads, 0.3, 0.5, max_det=100, cand_cap=64, check=False)
    cnt = r["count"].cpu()
    assert int(r["overflow"].item()) == 1 and int(cnt[0]) == -1 and int(cnt[1]) >= 0 and int(cnt[2]) == -1
    r = dh.detect_thresholded(heads, 0.3, 0.5, max_det=100, cand_cap=64, check=True)  # unfused GPU path
    check_against_oracle(dh, O, heads, r, 0.3, 0.5, 100)


def test_dense_detect_owned_workspace_stays_clean_over_calls(det, O):
    """DenseDetectWorkspace (gate bit 1: counters zeroed once, reset by the NMS CTAs, overflow word cleared by the select
    kernel): repeated calls on different batches -- an overflowing one in between -- give the oracle's detections, and
    the counter block is zero after every call."""
    C = 80
    dh = det.DenseAnchorHead(STRIDES, WH, C)
    ws = det.DenseDetectWorkspace(3, 1024, "cuda:0")
    nctr = det._native.fn("det_dense_detect_counter_bytes")(3)
    batches = [[h.cuda() for h in make_heads(3, 256, C, 40 + k, bias)] for k, bias in enumerate((-4.0, 0.0, -3.5, -4.0))]
    out = None
    for k, heads in enumerate(batches):
        out = dh.detect_thresholded(heads, 0.1, 0.5, max_det=300, cand_cap=1024, gate=bool(k & 1), check=False, out=out,
                                    workspace=ws)
        assert int(ws.buf[:nctr].view(torch.int32).abs().sum()) == 0
        if k == 1:  # bias 0: far more than 1024 candidates per image
            assert int(out["overflow"].item()) == 1 and (out["count"].cpu() == -1).all()
        else:
            assert int(out["overflow"].item()) == 0  # cleared again by the call after the overflowing one
            check_against_oracle(dh, O, heads, out, 0.1, 0.5, 300)
    with pytest.raises(ValueError):
        dh.detect_thresholded(batches[0], 0.1, 0.5, cand_cap=2048, workspace=ws)


def test_dense_detect_levels_outside_the_fused_limits(det, O):
    """5x5 level (25 positions, not a multiple of 4): detect_thresholded takes the unfused GPU path."""
    C = 10
    dh = det.DenseAnchorHead(STRIDES, WH, C)
    heads = [h.cuda() for h in make_heads(2, 160, C, 4, -2.0)]
    assert not dh.fused_ok(heads)
    r = dh.detect_thresholded(heads, 0.1, 0.5, max_det=50)
    check_against_oracle(dh, O, heads, r, 0.1, 0.5, 50)
    lv, hcs, n, na, _ = dh._levels(heads)
    import ctypes
    import det_b200._native as N
    with pytest.raises(N.DetError, match="status -2"):
        ws = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
        o = torch.empty(2 * 50 * 4, dtype=torch.float32, device="cuda")
        N.call("det_dense_detect", ctypes.cast(lv, ctypes.c_void_p), 3, n, na, C, 4.0, 0.1, 0.5, 0, 1, 1024, 50,
               N.ptr(o), N.ptr(o), N.ptr(o), N.ptr(o), N.ptr(o), None, N.ptr(ws), ws.numel(), N.stream())


def test_dense_detect_full_size_properties(det):
    """BASELINE configs[3] at full size (25 200 anchors x 80 classes, 8 images): size-independent properties --
    gated == ungated == unfused path, scores descending and above the threshold, rows unique, kept boxes of one class
    pairwise below the IoU threshold."""
    C = 80
    dh = det.DenseAnchorHead(STRIDES, WH, C)
    heads = [h.cuda() for h in make_heads(8, 640, C, 3, -4.0)]
    thr = 0.1
    a = dh.detect_thresholded(heads, thr, 0.5, max_det=300, cand_cap=2048, gate=True, check=False)
    b = dh.detect_thresholded(heads, thr, 0.5, max_det=300, cand_cap=2048, gate=False, check=False)
    c = dh._detect_thresholded_unfused(heads, thr, 0.5, 300, 0)
    assert int(a["overflow"].item()) == 0
    cnt = a["count"].cpu()
    assert torch.equal(cnt, b["count"].cpu()) and torch.equal(cnt, c["count"].cpu())
    assert int(cnt.min()) > 50
    for i in range(8):
        k = int(cnt[i])
        for key in ("idx", "boxes", "scores", "classes"):
            assert torch.equal(a[key][i, :k], b[key][i, :k]) and torch.equal(a[key][i, :k], c[key][i, :k]), key
        s = a["scores"][i, :k]
        assert bool((s[:-1] >= s[1:]).all()) and bool((s > thr).all())
        assert a["idx"][i, :k].unique().numel() == k
        bx, cl = a["boxes"][i, :k], a["classes"][i, :k]
        iou = det.pairwise_iou(det.Boxes(bx), det.Boxes(bx))
        same = (cl[:, None] == cl[None, :]) & ~torch.eye(k, dtype=torch.bool, device=bx.device)
        assert float((iou * same).max()) <= 0.5 + 1e-6


def test_threshold_compact_and_gather_kernels(det):
    """det_threshold_compact keeps the candidates in row order (torch.nonzero order), det_gather_detections maps kept
    candidate indices back to rows; NaN scores never pass; a cap below the candidate count truncates in row order."""
    import det_b200._native as N
    g = gen(21)
    n, R = 5, 3001
    boxes = torch.rand(n, R, 4, generator=g).cuda()
    scores = torch.rand(n, R, generator=g)
    scores[1] = 0.0                      # an image without candidates
    scores[2, ::7] = float("nan")
    scores = scores.cuda()
    classes = torch.randint(0, 80, (n, R), generator=g).cuda()
    for cap in (R, 100):
        rows = torch.full((n, cap), -1, dtype=torch.int64, device="cuda")
        cb = torch.zeros((n, cap, 4), device="cuda")
        cs = torch.zeros((n, cap), device="cuda")
        cc = torch.zeros((n, cap), dtype=torch.int64, device="cuda")
        cnt = torch.zeros((n,), dtype=torch.int32, device="cuda")
        N.call("det_threshold_compact", N.ptr(boxes), N.ptr(scores), N.ptr(classes), n, R, 0.6, cap, N.ptr(rows), N.ptr(cb),
               N.ptr(cs), N.ptr(cc), N.ptr(cnt), N.stream())
        for i in range(n):
            want = torch.nonzero(scores[i] > 0.6, as_tuple=True)[0][:cap]
            k = want.numel()
            assert int(cnt[i]) == k
            assert torch.equal(rows[i, :k], want) and torch.equal(cb[i, :k], boxes[i, want])
            assert torch.equal(cs[i, :k], scores[i, want]) and torch.equal(cc[i, :k], classes[i, want])
    # gather: keep = the first few candidates reversed
    max_det = 8
    keep = torch.zeros((n, max_det), dtype=torch.int64, device="cuda")
    kc = torch.minimum(cnt, torch.tensor(max_det, dtype=torch.int32, device="cuda")).to(torch.int32)
    for i in range(n):
        k = int(kc[i])
        keep[i, :k] = torch.arange(k - 1, -1, -1, device="cuda")
    out_idx = torch.full((n, max_det), -7, dtype=torch.int64, device="cuda")
    ob = torch.zeros((n, max_det, 4), device="cuda")
    os_ = torch.zeros((n, max_det), device="cuda")
    oc = torch.zeros((n, max_det), dtype=torch.int64, device="cuda")
    N.call("det_gather_detections", N.ptr(keep), N.ptr(kc), n, max_det, N.ptr(rows), N.ptr(cb), N.ptr(cs), N.ptr(cc), cap,
           N.ptr(out_idx), N.ptr(ob), N.ptr(os_), N.ptr(oc), N.stream())
    for i in range(n):
        k = int(kc[i])
        assert torch.equal(out_idx[i, :k], rows[i, :k].flip(0)) and torch.equal(ob[i, :k], cb[i, :k].flip(0))
        assert torch.equal(os_[i, :k], cs[i, :k].flip(0)) and torch.equal(oc[i, :k], cc[i, :k].flip(0))
        assert bool((out_idx[i, k:] == -7).all())


@pytest.mark.parametrize("seed", list(range(int(__import__("os").environ.get("DET_STRESS_SEEDS", "12")))))
def test_dense_detect_stress_random_heads(det, O, seed):
    """Random dense heads through det_dense_detect: image size, classes, objectness bias (from nothing to thousands of
    candidates per image), thresholds, max_det, candidate cap, gate on/off, quantised logits (ties), NMS branch: rows, counts,
    boxes, scores and classes bit-exact against the oracle on the decoded values; overflowing images take the exact route."""
    g = gen(11000 + seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    rf = lambda lo, hi: float(torch.rand(1, generator=g)) * (hi - lo) + lo
    image = 32 * ri(2, 8)
    C = [1, 3, 20, 80][ri(0, 3)]
    n = ri(1, 5)
    dh = det.DenseAnchorHead(STRIDES, WH, C)
    heads = make_heads(n, image, C, 500 + seed, rf(-5.0, 0.5))
    if ri(0, 2) == 0:
        heads = [(h * 4).round() / 4 for h in heads]
    heads = [h.cuda() for h in heads]
    thr, iou = rf(0.02, 0.5), rf(0.2, 0.8)
    max_det = [1, 30, 300, 1000][ri(0, 3)]
    cap = [256, 1024, 4096][ri(0, 2)]
    r = dh.detect_thresholded(heads, thr, iou, max_det=max_det, cand_cap=cap, gate=bool(ri(0, 1)), check=True)
    check_against_oracle(dh, O, heads, r, thr, iou, max_det)


def test_dense_detect_offset_trick_suppresses_across_categories(det, O):
    """<= 1000 candidates -> torchvision's coordinate-offset branch.  A top-left box of class 1 that sticks out of the frame
    by half its size, shifted by one span, overlaps the bottom-right box of class 0 that holds the largest coordinate:
    at a low IoU threshold the reference suppresses ACROSS categories, and so must we (either may be the better one)."""
    import math
    C = 4
    dh = det.DenseAnchorHead(STRIDES, WH, C)
    heads = make_heads(2, 640, C, 33, -7.0)
    v = heads[2].view(2, 3, 5 + C, 20, 20)
    for img, (obj_q, obj_p) in enumerate(((4.0, 9.0), (9.0, 4.0))):
        for (row, col, cls, txy, obj) in ((0, 0, 1, -20.0, obj_q), (19, 19, 0, 20.0, obj_p)):
            v[img, 2, :, row, col] = -10.0
            v[img, 2, 0, row, col] = txy
            v[img, 2, 1, row, col] = txy
            v[img, 2, 2, row, col] = math.log(600.0 / 373.0)
            v[img, 2, 3, row, col] = math.log(600.0 / 326.0)
            v[img, 2, 4, row, col] = obj
            v[img, 2, 5 + cls, row, col] = 10.0
    hg = [h.cuda() for h in heads]
    boxes, scores, classes = [t.cpu() for t in dh.decode(hg)]
    for gate in (True, False):
        r = dh.detect_thresholded(hg, 0.3, 0.1, max_det=300, cand_cap=1024, gate=gate, check=False)
        check_against_oracle(dh, O, hg, r, 0.3, 0.1, 300)
    for i in range(2):  # the case is what it claims to be: per-category NMS would keep one detection more
        cand = torch.nonzero(scores[i] > 0.3, as_tuple=True)[0]
        assert cand.numel() <= 1000
        per_cat = sum(int(O.nms(boxes[i][cand][classes[i][cand] == c], scores[i][cand][classes[i][cand] == c], 0.1).numel())
                      for c in range(C))
        assert int(r["count"][i]) == per_cat - 1
