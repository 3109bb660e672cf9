"""CPU: oracle/ref_torch.py pinned to golden vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_torch as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with np.load(os.path.join(GOLD, name + ".npz")) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


def test_pairwise_overlaps_bit_exact():
    d = load("pairwise")
    assert torch.equal(O.pairwise_iou(d["b1"], d["b2"]), d["iou"])
    assert torch.equal(O.pairwise_ioa(d["b1"], d["b2"]), d["ioa"])
    assert torch.equal(O.pairwise_intersection(d["b1"], d["b2"]), d["inter"])
    assert torch.equal(O.matched_boxlist_iou(d["b2"][:20], d["b1"]), d["miou"])


def test_codec_bit_exact_and_known_answer():
    d = load("codec")
    w = tuple(float(x) for x in d["weights"])
    assert torch.equal(O.apply_deltas(d["deltas"], d["boxes"], w), d["applied"])
    assert torch.equal(O.get_deltas(d["boxes"], d["tgt"], w), d["encoded"])
    ka = O.apply_deltas(d["ka_in"][1:2].float(), d["ka_in"][0:1].float(), w)
    assert torch.equal(ka, d["ka_out"])
    # SURVEY.md section 4 known answer
    torch.testing.assert_close(ka, torch.tensor([[-306.5, 5.8127, 318.5, 22.1873]]), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("off", [0.0, 0.5])
def test_anchors_bit_exact(off):
    d = load(f"anchors_off{int(off * 10)}")
    grids = [tuple(int(v) for v in g) for g in d["grids"]]
    cells = [O.cell_anchors(s, [0.5, 1.0, 2.0]) for s in ([32], [64], [128], [256], [512])]
    for c, want in zip(cells, d["cells"]):
        assert torch.equal(c, want)
    got = O.grid_anchors(grids, [4, 8, 16, 32, 64], cells, off)
    for i, a in enumerate(got):
        assert torch.equal(a, d[f"lvl{i}"])


def test_anchor_counts_fpn18_448():
    cells = [O.cell_anchors(s, [0.5, 1.0, 2.0]) for s in ([32], [64], [128], [256], [512])]
    got = O.grid_anchors([(448 // s, 448 // s) for s in (4, 8, 16, 32, 64)], [4, 8, 16, 32, 64], cells, 0.0)
    assert [len(a) for a in got] == [37632, 9408, 2352, 588, 147] and sum(len(a) for a in got) == 50127


def test_matcher_bit_exact():
    d = load("matcher")
    for name, (thr, lab, lq) in {"rpn": ([0.3, 0.7], [0, -1, 1], True), "roi": ([0.5], [0, 1], False),
                                 "rpn_nolq": ([0.3, 0.7], [0, -1, 1], False)}.items():
        i, l = O.match(d["q"], thr, lab, lq)
        assert l.dtype == torch.int8 and i.dtype == torch.int64
        assert torch.equal(i, d[name + "_idx"]) and torch.equal(l, d[name + "_lab"]), name
    i, l = O.match(torch.zeros(0, 5), [0.3, 0.7], [0, -1, 1], True)
    assert torch.equal(i, d["empty_idx"]) and torch.equal(l, d["empty_lab"])


def test_matcher_known_answers():
    # bucket rule low <= v < high
    _, l = O.match(torch.tensor([[0.3, 0.7, 0.29999998, 0.6999999]]), [0.3, 0.7], [0, -1, 1], False)
    assert l.tolist() == [-1, 1, 0, -1]
    # ties in the column max -> lowest gt index
    i, _ = O.match(torch.tensor([[0.4, 0.1], [0.4, 0.2]]), [0.3, 0.7], [0, -1, 1], False)
    assert i.tolist() == [0, 1]
    # a gt with best IoU 0 promotes every anchor with IoU 0 to it
    _, l = O.match(torch.tensor([[0.0, 0.0, 0.0], [0.1, 0.0, 0.2]]), [0.3, 0.7], [0, -1, 1], True)
    assert l.tolist() == [1, 1, 1]


@pytest.mark.parametrize("name", ["trick", "edge1000", "vanilla", "vanilla_thr07", "single"])
def test_batched_nms_bit_exact(name):
    d = load("batched_nms")
    keep = O.batched_nms(d[name + "_b"], d[name + "_s"], d[name + "_c"], float(d[name + "_thr"]))
    assert torch.equal(keep, d[name + "_keep"])


@pytest.mark.parametrize("name", ["eval", "train_small_topk"])
def test_find_top_rpn_proposals_bit_exact(name):
    d = load("proposals")
    pre, post, training = (int(v) for v in d[name + "_cfg"])
    props = [d[f"{name}_p{l}"].clone() for l in range(5)]
    logits = [d[f"{name}_l{l}"] for l in range(5)]
    sizes = [tuple(int(v) for v in s) for s in d[name + "_sizes"]]
    res = O.find_top_rpn_proposals(props, logits, sizes, 0.7, pre, post, 2.0, bool(training))
    for i, (b, s) in enumerate(res):
        assert torch.equal(b, d[f"{name}_boxes{i}"]) and torch.equal(s, d[f"{name}_logits{i}"])


def test_rpn_assignment_and_losses():
    d = load("rpn_train")
    gts = [d["gt0"], d["gt1"], d["gt2"]]
    labs, idxs = O.label_anchors(d["anchors"], gts)
    assert torch.equal(torch.stack(labs), d["pre_labels"]) and torch.equal(torch.stack(idxs), d["pre_idx"])
    # the sampled labels are a subset with the reference's counts
    for i in range(3):
        wp, wn = O.subsample_counts(int((labs[i] == 1).sum()), int((labs[i] == 0).sum()), 256, 0.5)
        assert int((d["labels"][i] == 1).sum()) == wp and int((d["labels"][i] == 0).sum()) == wn
        assert torch.equal(d["mboxes"][i], gts[i][idxs[i]])
    lg = d["logits"].clone().requires_grad_(True)
    dl = d["deltas"].clone().requires_grad_(True)
    out = O.rpn_losses(d["anchors"], lg, d["labels"], dl, d["mboxes"])
    (out["cls_loss"] + out["loc_loss"]).backward()
    assert torch.equal(out["cls_loss"].detach(), d["cls_loss"]) and torch.equal(out["loc_loss"].detach(), d["loc_loss"])
    assert torch.equal(lg.grad, d["grad_logits"]) and torch.equal(dl.grad, d["grad_deltas"])


def test_roi_matching():
    d = load("roi_match")
    allp, mi, ml, cls = O.roi_label_proposals(d["props"], d["gt"], d["gt_classes"], 80)
    assert torch.equal(mi, d["matched"]) and torch.equal(ml, d["labels"]) and torch.equal(cls, d["classes"])
