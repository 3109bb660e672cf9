"""GPU parity: ROI-head matching / sampling (reference python/src/models/roi.py:68-193)."""
import os

import numpy as np
import pytest
import torch

from tests.util import gen, rand_boxes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def det():
    import det_b200
    return det_b200


@pytest.fixture(scope="module")
def O():
    from oracle import ref_torch
    return ref_torch


def _heads(det, num_classes=80, bs=512, frac=0.25, append=True):
    return det.ROIHeads(num_classes, bs, frac, det.Matcher([0.5], [0, 1], allow_low_quality_matches=False), append)


def test_label_proposals_matches_reference_golden(det, O):
    """tests/golden/roi_match.npz was produced by the unmodified reference (pairwise_iou -> Matcher -> class labels)."""
    with np.load(os.path.join(GOLD, "roi_match.npz")) as z:
        d = {k: torch.from_numpy(z[k]) for k in z.files}
    h = _heads(det)
    allp = torch.cat([d["props"], d["gt"]], 0)  # the golden includes the appended gt boxes
    mi, ml, cls = h.label_proposals(allp.cuda(), d["gt"].cuda(), d["gt_classes"].cuda())
    assert torch.equal(mi.cpu(), d["matched"]) and torch.equal(ml.cpu(), d["labels"]) and torch.equal(cls.cpu(), d["classes"])


@pytest.mark.parametrize("P,G,seed", [(2000, 12, 0), (300, 1, 1), (64, 40, 2), (1000, 0, 3)])
def test_label_proposals_matches_oracle(det, O, P, G, seed):
    g = gen(seed)
    props, gtb = rand_boxes(P, 512.0, g, 0.3), rand_boxes(G, 512.0, g, 0.3)
    if G:
        props[: min(P, G)] = gtb[: min(P, G)] + torch.rand(min(P, G), 4, generator=g) * 6 - 3  # some real positives
    gtc = torch.randint(0, 80, (G,), generator=g)
    h = _heads(det)
    mi, ml, cls = h.label_proposals(props.cuda(), gtb.cuda(), gtc.cuda())
    _, wi, wl, wc = O.roi_label_proposals(props, gtb, gtc, 80, append_gt=False)
    assert torch.equal(mi.cpu(), wi) and torch.equal(ml.cpu(), wl) and torch.equal(cls.cpu(), wc)


def test_label_and_sample_proposals_contract(det, O):
    """Counts, class labels, appended gt, copied gt_* fields and the no-gt image, as the reference defines them."""
    g = gen(7)
    h = _heads(det, num_classes=20, bs=128, frac=0.25)
    proposals, targets, raw = [], [], []
    for i, (P, G) in enumerate([(500, 6), (40, 3), (300, 0)]):
        pb, gb = rand_boxes(P, 256.0, g, 0.4), rand_boxes(G, 256.0, g, 0.4)
        if G:
            pb[:G] = gb + torch.rand(G, 4, generator=g) * 4 - 2
        gc = torch.randint(0, 20, (G,), generator=g)
        p = det.Instances((256, 256))
        p.proposal_boxes = det.Boxes(pb.cuda())
        p.objectness_logits = torch.randn(P, generator=g).cuda()
        t = det.Instances((256, 256))
        t.gt_boxes = det.Boxes(gb.cuda())
        t.gt_classes = gc.cuda()
        proposals.append(p)
        targets.append(t)
        raw.append((pb, gb, gc))
    torch.manual_seed(0)
    out = h.label_and_sample_proposals(proposals, targets)
    assert len(out) == 3
    for (pb, gb, gc), res in zip(raw, out):
        P, G = pb.shape[0], gb.shape[0]
        allp, wi, wl, wc = O.roi_label_proposals(pb, gb, gc, 20)  # gt appended
        cls = res.gt_classes.cpu()
        npos_av, nneg_av = int(((wc != -1) & (wc != 20)).sum()), int((wc == 20).sum())
        want_pos, want_neg = O.subsample_counts(npos_av, nneg_av, 128, 0.25)
        assert int((cls != 20).sum()) == want_pos and int((cls == 20).sum()) == want_neg
        # every sampled proposal is one of the P + G candidates and carries that candidate's label / matched gt
        boxes = res.proposal_boxes.tensor.cpu()
        for j in range(len(res)):
            hit = (allp == boxes[j]).all(dim=1).nonzero()[:, 0]
            assert hit.numel() >= 1
            assert int(cls[j]) in {int(wc[k]) for k in hit}
        if G:
            assert res.has("gt_boxes")
            gtb_sel = res.gt_boxes.tensor.cpu()
            fg = cls != 20
            # a foreground sample's gt box has IoU >= 0.5 with it and its class is the gt's class
            iou = O.pairwise_iou(gtb_sel[fg], boxes[fg]).diagonal()
            assert bool((iou >= 0.5).all())
            assert int(fg.sum()) >= min(G, 32)  # the appended gt boxes are positives themselves
        else:
            assert not res.has("gt_boxes") and bool((cls == 20).all())
    assert h.last_num_fg_samples is not None and h.last_num_bg_samples is not None
