"""CPU, build container only: randomized side-by-side of oracle/ref_torch.py and the unmodified reference functions."""
import pytest
import torch

from oracle import ref_loader, ref_torch as O
from tests.util import gen, rand_boxes, distinct_scores

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")


@pytest.fixture(scope="module")
def R():
    return ref_loader.load()


@pytest.mark.parametrize("seed", range(5))
def test_overlaps_and_codec(R, seed):
    g = gen(seed)
    b1, b2 = rand_boxes(17 + seed, 500.0, g), rand_boxes(400 + 13 * seed, 500.0, g)
    assert torch.equal(O.pairwise_iou(b1, b2), R.pairwise_iou(R.Boxes(b1), R.Boxes(b2)))
    assert torch.equal(O.pairwise_ioa(b1, b2), R.pairwise_ioa(R.Boxes(b1), R.Boxes(b2)))
    t = R.Box2BoxTransform((10.0, 10.0, 5.0, 5.0), O.DEFAULT_SCALE_CLAMP)
    d = torch.randn(b2.shape[0], 4, generator=g) * 4
    assert torch.equal(O.apply_deltas(d, b2, (10.0, 10.0, 5.0, 5.0)), t.apply_deltas(d, b2))
    tg = rand_boxes(b2.shape[0], 500.0, g)
    assert torch.equal(O.get_deltas(b2, tg, (10.0, 10.0, 5.0, 5.0)), t.get_deltas(b2, tg))


@pytest.mark.parametrize("seed", range(12))
def test_batched_nms(R, seed):
    g = gen(100 + seed)
    n = [50, 999, 1000, 1001, 2000, 3500][seed % 6]
    b = rand_boxes(n, 400.0, g)
    if seed % 2:
        b = b.round()
    s, c = distinct_scores(n, g), torch.randint(0, [1, 6, 80][seed % 3], (n,), generator=g)
    thr = [0.5, 0.7, 0.3][seed % 3]
    assert torch.equal(O.batched_nms(b, s, c, thr), R.batched_nms(b, s, c, thr))


def test_batched_nms_over_40000_uses_reference_loop(R):
    g = gen(7)
    n = 40500
    b, s, c = rand_boxes(n, 3000.0, g, 0.02), distinct_scores(n, g), torch.randint(0, 30, (n,), generator=g)
    assert torch.equal(O.batched_nms(b, s, c, 0.5), R.batched_nms(b, s, c, 0.5))


@pytest.mark.parametrize("seed", range(4))
def test_matcher(R, seed):
    g = gen(200 + seed)
    q = O.pairwise_iou(rand_boxes(1 + 7 * seed, 300.0, g), rand_boxes(900, 300.0, g, 0.3))
    for thr, lab, lq in (([0.3, 0.7], [0, -1, 1], True), ([0.5], [0, 1], False)):
        a = O.match(q, thr, lab, lq)
        b = R.Matcher(thr, lab, lq)(q.clone())
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_find_top_rpn_proposals(R):
    g = gen(300)
    props, logits = [], []
    ks = (3000, 700, 150, 40, 9)
    pool = distinct_scores(2 * sum(ks), g)  # tie-free across levels too (the reference's final sort is unstable)
    for k in ks:
        props.append(rand_boxes(2 * k, 260.0, g, 0.4).view(2, k, 4) - 20.0)
        logits.append(pool[:2 * k].view(2, k))
        pool = pool[2 * k:]
    sizes = [(200, 220), (180, 200)]
    a = O.find_top_rpn_proposals([p.clone() for p in props], logits, sizes, 0.7, 1000, 300, 1.0, False)
    b = R.find_top_rpn_proposals([p.clone() for p in props], logits, sizes, 0.7, 1000, 300, 1.0, False)
    for (ab, as_), bi in zip(a, b):
        assert torch.equal(ab, bi.proposal_boxes.tensor) and torch.equal(as_, bi.objectness_logits)


@pytest.mark.parametrize("seed", range(4))
def test_dense_select_nms_is_the_reference_batched_nms_on_the_thresholded_rows(R, seed):
    """oracle.dense_select_nms (the spec of det_dense_detect's selection stage) = rows with score > thr, in row order,
    through the REFERENCE's batched_nms, cut at max_det -- including negative coordinates (boxes sticking out of the
    frame), where torchvision's coordinate-offset trick can suppress across categories."""
    g = gen(400 + seed)
    n = [300, 1500, 5000, 900][seed]
    b = rand_boxes(n, 640.0, g, 0.3) - 60.0
    s = torch.rand(n, generator=g)
    if seed % 2:
        s = (s * 64).round() / 64  # ties: row order decides
    c = torch.randint(0, 80, (n,), generator=g)
    thr, iou, max_det = 0.4, 0.5, 200
    rows, bb, ss, cc = O.dense_select_nms(b, s, c, thr, iou, max_det)
    cand = torch.nonzero(s > thr, as_tuple=True)[0]
    keep = R.batched_nms(b[cand], s[cand], c[cand], iou)[:max_det]
    assert torch.equal(rows, cand[keep])
    assert torch.equal(bb, b[cand][keep]) and torch.equal(ss, s[cand][keep]) and torch.equal(cc, c[cand][keep])
