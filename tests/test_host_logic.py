"""CPU: host-side mirror of the reference interface (containers, constructors, error behaviour) -- no kernels run."""
import math

import pytest
import torch

import det_b200 as det
from oracle import ref_torch as O


def test_boxes_container():
    b = det.Boxes(torch.tensor([[0.0, 0.0, 10.0, 5.0], [2.0, 2.0, 2.0, 9.0]]))
    assert len(b) == 2 and b.area().tolist() == [50.0, 0.0]
    assert b.nonempty().tolist() == [True, False]
    assert len(det.Boxes(torch.empty(0))) == 0 and det.Boxes(torch.empty(0)).tensor.shape == (0, 4)
    c = b.clone()
    c.clip((4, 8))
    assert c.tensor.tolist() == [[0.0, 0.0, 8.0, 4.0], [2.0, 2.0, 2.0, 4.0]]
    assert len(det.Boxes.cat([b, c])) == 4 and len(b[1]) == 1 and len(b[torch.tensor([True, False])]) == 1
    with pytest.raises(AssertionError):
        det.Boxes(torch.zeros(3, 5))
    with pytest.raises(AssertionError):
        det.Boxes(torch.tensor([[float("nan"), 0, 1, 1]])).clip((4, 4))


def test_instances_container():
    i = det.Instances((10, 20))
    i.proposal_boxes = det.Boxes(torch.zeros(3, 4))
    i.objectness_logits = torch.arange(3.0)
    assert len(i) == 3 and i.image_size == (10, 20) and i.has("objectness_logits")
    assert len(i[torch.tensor([0, 2])]) == 2 and i[1].objectness_logits.tolist() == [1.0]
    with pytest.raises(AssertionError):
        i.bad = torch.zeros(5)
    j = det.Instances.cat([i, i])
    assert len(j) == 6 and isinstance(j.proposal_boxes, det.Boxes)
    with pytest.raises(AttributeError):
        _ = i.missing


def test_matcher_constructor_checks():
    m = det.Matcher([0.3, 0.7], [0, -1, 1], True)
    assert m.thresholds == [-float("inf"), 0.3, 0.7, float("inf")] and m.labels == [0, -1, 1]
    with pytest.raises(AssertionError):
        det.Matcher([0.7, 0.3], [0, -1, 1])
    with pytest.raises(AssertionError):
        det.Matcher([0.5], [0, 2])
    with pytest.raises(AssertionError):
        det.Matcher([0.5], [0, 1, 1])


def test_cell_anchors_match_oracle_and_known_values():
    c = det.generate_cell_anchors([32], [0.5, 1.0, 2.0]).float()
    assert torch.equal(c, O.cell_anchors([32], [0.5, 1.0, 2.0]))
    torch.testing.assert_close(c[0], torch.tensor([-22.6274, -11.3137, 22.6274, 11.3137]), rtol=1e-5, atol=1e-4)
    ag = det.AnchorGenerator([4, 8, 16, 32, 64])
    assert ag.num_anchors == [3] * 5
    with pytest.raises(AssertionError):
        det.AnchorGenerator([4], offset=1.0)


def test_defaults_follow_reference_config():
    rpn = det.RegionProposalNetwork([4, 8, 16, 32, 64])
    assert rpn.batch_size_per_image == 256 and rpn.positive_fraction == 0.5
    assert rpn.pre_nms_topk == (12000, 6000) and rpn.post_nms_topk == (2000, 1000) and rpn.nms_thresh == 0.7
    assert rpn.box2box_transform.weights == (1.0, 1.0, 1.0, 1.0)
    assert rpn.box2box_transform.scale_clamp == math.log(1000.0 / 16)
    assert rpn.pre_nms_topk[False] == 12000 and rpn.pre_nms_topk[True] == 6000  # indexed by `training` (rpn.py:324)


def test_cpu_tensors_raise_no_fallback():
    b = torch.zeros(2, 4)
    for fn in (lambda: det.pairwise_iou(b, b), lambda: det.batched_nms(b, torch.zeros(2), torch.zeros(2, dtype=torch.int64), 0.5),
               lambda: det.Box2BoxTransform().apply_deltas(b, b), lambda: det.Matcher([0.5], [0, 1])(torch.zeros(1, 2)),
               lambda: det.YoloGridHead().detect(torch.zeros(1, 7, 7, 30))):
        with pytest.raises(RuntimeError):
            fn()


def test_product_package_does_not_import_oracle():
    import os
    root = os.path.dirname(os.path.abspath(det.__file__))
    for f in os.listdir(root):
        if f.endswith(".py"):
            src = open(os.path.join(root, f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f


def test_shard_range_partitions_batch():
    for n in (1, 7, 1024):
        for w in (1, 2, 3, 8):
            spans = [det.dist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_subsample_labels_mirror_consumes_rng_like_the_reference():
    """det.subsample_labels (reference python/src/utils.py:34-76) is plain torch: under the same seed it returns the
    oracle's -- hence the reference's -- indices, on CPU tensors too (it launches no kernel of ours)."""
    g = torch.Generator().manual_seed(3)
    labels = torch.randint(-1, 5, (4000,), generator=g)
    for num, frac, bg in [(256, 0.5, 0), (512, 0.25, 4), (16, 0.5, 2), (100000, 0.5, 0)]:
        torch.manual_seed(11)
        p1, n1 = det.subsample_labels(labels, num, frac, bg)
        torch.manual_seed(11)
        p2, n2 = O.subsample_labels(labels, num, frac, bg)
        assert torch.equal(p1, p2) and torch.equal(n1, n2)
        wp, wn = O.subsample_counts(int(((labels != -1) & (labels != bg)).sum()), int((labels == bg).sum()), num, frac)
        assert (p1.numel(), n1.numel()) == (wp, wn)
    e = torch.empty(0, dtype=torch.int64)
    p, n = det.subsample_labels(e, 8, 0.5, 0)
    assert p.numel() == 0 and n.numel() == 0


def test_roi_heads_constructor_mirrors_reference():
    m = det.Matcher([0.5], [0, 1], allow_low_quality_matches=False)
    h = det.ROIHeads(num_classes=80, batch_size_per_image=512, positive_fraction=0.25, proposal_matcher=m)
    assert (h.num_classes, h.batch_size_per_image, h.positive_fraction, h.proposal_append_gt) == (80, 512, 0.25, True)

    class _MC:
        thresholds, labels, allow_low_quality_matches = [0.5], [0, 1], False

    class _RC:
        num_classes, batch_size_per_image, positive_fraction, proposal_append_gt = 3, 64, 0.5, False
        proposal_matcher = _MC()

    b = det.ROIHeads.build(_RC())
    assert (b.num_classes, b.batch_size_per_image, b.proposal_append_gt) == (3, 64, False)
