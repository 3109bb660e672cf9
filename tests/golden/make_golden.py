#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (/root/reference) -- run in the build container only.

    python tests/golden/make_golden.py

Every fixture stores the inputs and the outputs the reference's own functions produced for them on CPU
(torch 2.11.0, torchvision 0.26.0).  tests/test_oracle_golden.py then pins oracle/ref_torch.py to these files on any
machine (the GPU box has no /root/reference).  Scores are tie-free because the reference's descending sorts are
unstable (see oracle/ref_torch.py docstring).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from tests.util import gen, rand_boxes, distinct_scores  # noqa: E402

R = ref_loader.load()
STRIDES = [4, 8, 16, 32, 64]


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        out[k] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path)} bytes")


def main():
    torch.manual_seed(0)
    g = gen(1234)
    # ---- pairwise overlaps (structures/boxes.py)
    b1, b2 = rand_boxes(20, 448.0, g), rand_boxes(300, 448.0, g)
    b2[::9] = b1[3]
    b2[4, 2:] = b2[4, :2]
    save("pairwise", b1=b1, b2=b2, iou=R.pairwise_iou(R.Boxes(b1), R.Boxes(b2)),
         ioa=R.pairwise_ioa(R.Boxes(b1), R.Boxes(b2)), inter=R.pairwise_intersection(R.Boxes(b1), R.Boxes(b2)),
         miou=R.matched_boxlist_iou(R.Boxes(b2[:20]), R.Boxes(b1)))
    # ---- box codec (components/box_regression.py)
    w = (10.0, 10.0, 5.0, 5.0)
    t = R.Box2BoxTransform(w, float(np.log(1000.0 / 16)))
    boxes, tgt = rand_boxes(200, 448.0, g), rand_boxes(200, 448.0, g)
    deltas = torch.randn(200, 8, generator=g) * 3
    save("codec", weights=np.array(w), boxes=boxes, tgt=tgt, deltas=deltas, applied=t.apply_deltas(deltas, boxes),
         encoded=t.get_deltas(boxes, tgt),
         ka_in=np.array([[0.0, 0.0, 10.0, 20.0], [1.0, 2.0, 30.0, -1.0]]),
         ka_out=t.apply_deltas(torch.tensor([[1.0, 2.0, 30.0, -1.0]]), torch.tensor([[0.0, 0.0, 10.0, 20.0]])))
    # ---- anchors (modules/anchor_generators.py)
    for off in (0.0, 0.5):
        ag = R.AnchorGenerator(STRIDES, [[32], [64], [128], [256], [512]], [[0.5, 1.0, 2.0]], off, 4)
        feats = [torch.zeros(1, 1, 96 // s if 96 // s else 1, 128 // s if 128 // s else 1) for s in STRIDES]
        anc = ag(feats)
        save(f"anchors_off{int(off * 10)}", grids=np.array([f.shape[-2:] for f in feats]),
             cells=torch.stack(list(ag.cell_anchors._buffers.values())), **{f"lvl{i}": a.tensor for i, a in enumerate(anc)})
    # ---- Matcher (components/matcher.py)
    gt, an = rand_boxes(11, 300.0, g), rand_boxes(700, 300.0, g, 0.3)
    q = R.pairwise_iou(R.Boxes(gt), R.Boxes(an))
    q[5] = 0.0
    q[:, 17] = q[2, 17]
    out = {"q": q}
    for name, (thr, lab, lq) in {"rpn": ([0.3, 0.7], [0, -1, 1], True), "roi": ([0.5], [0, 1], False),
                                 "rpn_nolq": ([0.3, 0.7], [0, -1, 1], False)}.items():
        i, l = R.Matcher(thr, lab, lq)(q.clone())
        out[name + "_idx"], out[name + "_lab"] = i, l
    i, l = R.Matcher([0.3, 0.7], [0, -1, 1], True)(torch.zeros(0, 5))
    out["empty_idx"], out["empty_lab"] = i, l
    save("matcher", **out)
    # ---- batched_nms (utils.py -> torchvision)
    out = {}
    for name, (n, ncat, thr) in {"trick": (300, 5, 0.5), "edge1000": (1000, 20, 0.5), "vanilla": (1500, 7, 0.5),
                                 "vanilla_thr07": (2500, 80, 0.7), "single": (800, 1, 0.3)}.items():
        b, s, c = rand_boxes(n, 300.0, g), distinct_scores(n, g), torch.randint(0, ncat, (n,), generator=g)
        out[name + "_b"], out[name + "_s"], out[name + "_c"] = b, s, c
        out[name + "_thr"] = np.array(thr)
        out[name + "_keep"] = R.batched_nms(b, s, c, thr)
    save("batched_nms", **out)
    # ---- find_top_rpn_proposals (models/utils.py)
    img = 96
    out = {}
    for name, (pre, post, training) in {"eval": (12000, 2000, False), "train_small_topk": (60, 25, True)}.items():
        props, logits = [], []
        for s in STRIDES:
            k = max(img // s, 1) ** 2 * 3
            bb = rand_boxes(2 * k, float(img) * 1.2, g, 0.4).view(2, k, 4) - 10.0
            props.append(bb)
            logits.append(distinct_scores(2 * k, g).view(2, k) * 8 - 4)
        sizes = [(img, img), (img - 11, img - 30)]
        res = R.find_top_rpn_proposals([p.clone() for p in props], logits, sizes, 0.7, pre, post, 2.0, training)
        for l in range(len(STRIDES)):
            out[f"{name}_p{l}"], out[f"{name}_l{l}"] = props[l], logits[l]
        out[name + "_sizes"] = np.array(sizes)
        out[name + "_cfg"] = np.array([pre, post, int(training)])
        for i, r_i in enumerate(res):
            out[f"{name}_boxes{i}"] = r_i.proposal_boxes.tensor
            out[f"{name}_logits{i}"] = r_i.objectness_logits
    save("proposals", **out)
    # ---- label assignment + losses through the reference RPN object (models/rpn.py)
    from python.src.utils import ShapeSpec
    conf = R.config.RegionProposalNetworkConf()
    shapes = {f: ShapeSpec(64, 64, 3, s, 1, 1) for f, s in zip(conf.in_features, STRIDES)}
    rpn = R.RegionProposalNetwork.build(conf, shapes)
    img = 128
    feats = [torch.zeros(1, 64, img // s, img // s) for s in STRIDES]
    anchors = rpn.anchor_generator(feats)
    at = R.Boxes.cat(anchors).tensor
    n = 3
    gts = [rand_boxes(k, float(img), g, 0.5).clamp(max=float(img)) for k in (4, 1, 9)]
    pre_labels, pre_idx = [], []
    for gt_i in gts:  # rpn.py:167-168 without the random subsample
        qq = R.pairwise_iou(R.Boxes(gt_i), R.Boxes(at))
        i, l = rpn.anchor_matcher(qq)
        pre_labels.append(l)
        pre_idx.append(i)
    insts = []
    for gt_i in gts:
        inst = R.Instances((img, img))
        inst.gt_boxes = R.Boxes(gt_i)
        insts.append(inst)
    torch.manual_seed(7)
    labels, mboxes = rpn.label_and_sample_anchors(anchors, insts)
    lsz = [len(a) for a in anchors]
    logits = torch.randn(n, at.shape[0], generator=g).requires_grad_(True)
    deltas = (torch.randn(n, at.shape[0], 4, generator=g) * 0.5).requires_grad_(True)
    losses = rpn.losses(anchors, list(torch.split(logits, lsz, dim=1)), labels, list(torch.split(deltas, lsz, dim=1)),
                        mboxes)
    (losses["cls_loss"] + losses["loc_loss"]).backward()
    save("rpn_train", anchors=at, level_sizes=np.array(lsz), gt0=gts[0], gt1=gts[1], gt2=gts[2],
         pre_labels=torch.stack(pre_labels), pre_idx=torch.stack(pre_idx), labels=torch.stack(labels),
         mboxes=torch.stack(mboxes), logits=logits, deltas=deltas, cls_loss=losses["cls_loss"],
         loc_loss=losses["loc_loss"], grad_logits=logits.grad, grad_deltas=deltas.grad)
    # ---- ROI matching (models/roi.py), deterministic half
    props = rand_boxes(150, float(img), g, 0.5)
    gtb, gtc = gts[2], torch.randint(0, 80, (9,), generator=g)
    allp = torch.cat([props, gtb])
    qq = R.pairwise_iou(R.Boxes(gtb), R.Boxes(allp))
    mi, ml = R.Matcher([0.5], [0, 1], False)(qq)
    cls = gtc[mi].clone()
    cls[ml == 0] = 80
    cls[ml == -1] = -1
    save("roi_match", props=props, gt=gtb, gt_classes=gtc, matched=mi, labels=ml, classes=cls)


if __name__ == "__main__":
    main()
