"""GPU parity: box math kernels vs the CPU oracle (bit-exact for IoU family; 1e-5 rel for exp/log codecs)."""
import pytest
import torch

from tests.util import gen, rand_boxes, assert_boxes_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def det():
    import det_b200
    return det_b200


@pytest.fixture(scope="module")
def O():
    from oracle import ref_torch
    return ref_torch


@pytest.mark.parametrize("n,m", [(1, 1), (3, 5), (32, 50127), (37, 1000), (33, 1027), (700, 4096), (1, 4), (5, 0), (0, 7)])
def test_pairwise_iou_bit_exact(det, O, n, m):
    g = gen(n * 1000 + m)
    b1, b2 = rand_boxes(n, g=g), rand_boxes(m, g=g)
    if n and m:
        b2[::7] = b1[0]            # identical boxes -> IoU 1
        b2[1::11, 2:] = b2[1::11, :2]  # empty boxes
    for name in ("pairwise_iou", "pairwise_ioa", "pairwise_intersection"):
        want = getattr(O, name)(b1, b2)
        got = getattr(det, name)(det.Boxes(b1.cuda()), det.Boxes(b2.cuda())).cpu()
        assert got.shape == want.shape
        assert torch.equal(got, want), name


def test_pairwise_iou_nan_inf_and_touching(det, O):
    b1 = torch.tensor([[0, 0, 10, 10], [float("nan"), 0, 5, 5], [0, 0, float("inf"), 5], [5, 5, 5, 5],
                       [10, 0, 20, 10], [-float("inf"), -float("inf"), float("inf"), float("inf")]])
    b2 = torch.tensor([[0, 0, 10, 10], [10, 10, 20, 20], [2, 2, 3, float("nan")], [4, 4, 6, 6], [9, 9, 1, 1]])
    want = O.pairwise_iou(b1, b2)
    got = det.pairwise_iou(b1.cuda(), b2.cuda()).cpu()
    assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(want, nan=-7.0))


def test_matched_boxlist_iou(det, O):
    g = gen(5)
    b1, b2 = rand_boxes(3001, g=g), rand_boxes(3001, g=g)
    b2[::5] = b1[::5]
    want = O.matched_boxlist_iou(b1, b2)
    got = det.matched_boxlist_iou(det.Boxes(b1.cuda()), det.Boxes(b2.cuda())).cpu()
    assert torch.equal(got, want)


@pytest.mark.parametrize("weights", [(1.0, 1.0, 1.0, 1.0), (10.0, 10.0, 5.0, 5.0)])
@pytest.mark.parametrize("k", [1, 3])
def test_apply_deltas(det, O, weights, k):
    g = gen(11)
    boxes = rand_boxes(5003, g=g)
    deltas = torch.randn(5003, 4 * k, generator=g) * 3
    want = O.apply_deltas(deltas, boxes, weights)
    t = det.Box2BoxTransform(weights, O.DEFAULT_SCALE_CLAMP)
    got = t.apply_deltas(deltas.cuda(), boxes.cuda()).cpu()
    # fp32, 1e-5 relative (north star); exp differs by <= 2 ulp between CUDA and the CPU vector library
    assert_boxes_close(got.reshape(-1, 4), want.reshape(-1, 4), rtol=1e-5)


@pytest.mark.parametrize("k", [1, 3])
def test_apply_deltas_is_differentiable_like_the_reference(det, O, k):
    """box_regression.py:75-115 is plain torch: autograd flows to the deltas (and the boxes).  The drop-in's backward
    kernel (det_apply_deltas_backward) must give the same gradients, including the clamp(max=) mask."""
    g = gen(13)
    w = (10.0, 10.0, 5.0, 5.0)
    boxes = rand_boxes(777, g=g)
    deltas = torch.randn(777, 4 * k, generator=g) * 3
    deltas[5, 2] = 40.0  # dw = 8 > scale_clamp: clamped, zero gradient
    up = torch.randn(777, 4 * k, generator=g)
    d0, b0 = deltas.clone().requires_grad_(True), boxes.clone().requires_grad_(True)
    (O.apply_deltas(d0, b0, w) * up).sum().backward()
    d1, b1 = deltas.cuda().requires_grad_(True), boxes.cuda().requires_grad_(True)
    out = det.Box2BoxTransform(w, O.DEFAULT_SCALE_CLAMP).apply_deltas(d1, b1)
    assert out.requires_grad
    (out * up.cuda()).sum().backward()
    assert float(d1.grad[5, 2]) == 0.0
    torch.testing.assert_close(d1.grad.cpu(), d0.grad, rtol=2e-5, atol=1e-5)
    torch.testing.assert_close(b1.grad.cpu(), b0.grad, rtol=2e-5, atol=1e-4)
    # no graph is built when nothing requires grad
    assert not det.Box2BoxTransform(w).apply_deltas(deltas.cuda(), boxes.cuda()).requires_grad


def test_apply_deltas_known_answer(det):
    # weights (10,10,5,5), box [0,0,10,20], delta [1,2,30,-1] -> dw clamped to log(1000/16)
    import math
    t = det.Box2BoxTransform((10.0, 10.0, 5.0, 5.0), math.log(1000.0 / 16))
    out = t.apply_deltas(torch.tensor([[1.0, 2.0, 30.0, -1.0]]).cuda(), torch.tensor([[0.0, 0.0, 10.0, 20.0]]).cuda())
    torch.testing.assert_close(out.cpu(), torch.tensor([[-306.5, 5.8127, 318.5, 22.1873]]), rtol=1e-5, atol=1e-4)


def test_get_deltas(det, O):
    g = gen(12)
    src, tgt = rand_boxes(4001, g=g), rand_boxes(4001, g=g)
    w = (10.0, 10.0, 5.0, 5.0)
    want = O.get_deltas(src, tgt, w)
    got = det.Box2BoxTransform(w).get_deltas(src.cuda(), tgt.cuda()).cpu()
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)
    bad = src.clone()
    bad[5, 2] = bad[5, 0]
    with pytest.raises(AssertionError):
        det.Box2BoxTransform(w).get_deltas(bad.cuda(), tgt.cuda())


@pytest.mark.parametrize("offset", [0.0, 0.5])
def test_grid_anchors_bit_exact(det, O, offset):
    strides = [4, 8, 16, 32, 64]
    sizes = [[32], [64], [128], [256], [512]]
    ratios = [[0.5, 1.0, 2.0]]
    grids = [(448 // s, 448 // s) for s in strides]
    cells = [O.cell_anchors(s, ratios[0]) for s in sizes]
    want = O.grid_anchors(grids, strides, cells, offset)
    ag = det.AnchorGenerator(strides, sizes, ratios, offset)
    feats = [torch.zeros(1, 1, h, w, device="cuda") for h, w in grids]
    got = ag(feats)
    assert [len(a) for a in got] == [37632, 9408, 2352, 588, 147]
    for a, b in zip(got, want):
        assert torch.equal(a.tensor.cpu(), b)


def test_cpu_tensor_raises(det):
    with pytest.raises(RuntimeError):
        det.pairwise_iou(torch.zeros(2, 4), torch.zeros(2, 4))
