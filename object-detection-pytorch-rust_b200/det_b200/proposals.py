"""RPN proposal selection -- drop-in for ``find_top_rpn_proposals`` (reference python/src/models/utils.py:9-109) and
``add_ground_truth_to_proposals`` (:111-155).  Arithmetic: det_rpn_proposals (csrc/proposals.cu)."""
import ctypes
import math
from typing import List, Sequence, Tuple

import torch

from . import _native as N
from .structures import Boxes, Instances


def rpn_proposals_batched(boxes: torch.Tensor, logits: torch.Tensor, level_sizes: Sequence[int],
                          image_sizes: torch.Tensor, nms_thresh: float, pre_nms_topk: int, post_nms_topk: int,
                          min_box_size: float, workspace: torch.Tensor = None):
    """Whole batch, no host synchronisation.  boxes (N,R,4), logits (N,R) with the levels concatenated along R,
    image_sizes (N,2) int32 device tensor of (h,w).  Returns (out_boxes (N,K,4), out_logits (N,K), counts (N) int32,
    nonfinite_flag (1) int32) with K = post_nms_topk."""
    N.require_cuda(boxes, logits, image_sizes)
    b, s = N.f32c(boxes), N.f32c(logits)
    n, r = s.shape
    assert b.shape == (n, r, 4) and sum(level_sizes) == r
    dev = b.device
    k = int(post_nms_topk)
    out_b = torch.empty((n, k, 4), dtype=torch.float32, device=dev)
    out_s = torch.empty((n, k), dtype=torch.float32, device=dev)
    cnt = torch.empty((n,), dtype=torch.int32, device=dev)
    flag = torch.zeros((1,), dtype=torch.int32, device=dev)
    if n == 0:
        return out_b, out_s, cnt, flag
    wsb = N.fn("det_rpn_proposals_workspace_bytes")(n, r)
    if workspace is None or workspace.numel() < wsb:
        workspace = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    lv = (ctypes.c_int64 * len(level_sizes))(*[int(x) for x in level_sizes])
    sizes = image_sizes.to(torch.int32).contiguous()
    with torch.cuda.device(dev):
        N.call("det_rpn_proposals", N.ptr(b), N.ptr(s), n, r, lv, len(level_sizes), N.ptr(sizes), float(nms_thresh),
               int(pre_nms_topk), k, float(min_box_size), N.ptr(out_b), N.ptr(out_s), N.ptr(cnt), N.ptr(flag),
               N.ptr(workspace), workspace.numel(), N.stream())
    return out_b, out_s, cnt, flag


def find_top_rpn_proposals(proposals: List[torch.Tensor], pred_objectness_logits: List[torch.Tensor],
                           image_sizes: List[Tuple[int, int]], nms_thresh: float, pre_nms_topk: int,
                           post_nms_topk: int, min_box_size: float, training: bool) -> List[Instances]:
    """Reference signature: L x (N,HiWiA,4), L x (N,HiWiA) -> N Instances {proposal_boxes, objectness_logits}."""
    level_sizes = [int(l.shape[1]) for l in pred_objectness_logits]
    boxes = proposals[0] if len(proposals) == 1 else torch.cat(proposals, dim=1)
    logits = pred_objectness_logits[0] if len(pred_objectness_logits) == 1 else torch.cat(pred_objectness_logits, dim=1)
    sizes = torch.tensor([[int(h), int(w)] for h, w in image_sizes], dtype=torch.int32).to(boxes.device)
    out_b, out_s, cnt, flag = rpn_proposals_batched(boxes, logits, level_sizes, sizes, nms_thresh, pre_nms_topk,
                                                    post_nms_topk, min_box_size)
    counts = cnt.tolist()  # the List[Instances] API has data-dependent lengths: one synchronisation for the batch
    if training and int(flag.item()):
        raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
    results = []
    for i, image_size in enumerate(image_sizes):
        res = Instances(image_size)
        res.proposal_boxes = Boxes(out_b[i, :counts[i]])
        res.objectness_logits = out_s[i, :counts[i]]
        results.append(res)
    return results


def add_ground_truth_to_proposals(gt_boxes: List[Boxes], proposals: List[Instances]) -> List[Instances]:
    """Append the gt boxes to each image's proposals with logit(1 - 1e-10) (reference models/utils.py:111-155)."""
    assert gt_boxes is not None
    assert len(proposals) == len(gt_boxes)
    if len(proposals) == 0:
        return proposals
    gt_logit_value = math.log((1.0 - 1e-10) / (1 - (1.0 - 1e-10)))
    out = []
    for gt_i, prop_i in zip(gt_boxes, proposals):
        dev = prop_i.objectness_logits.device
        rec = type(prop_i)  # duck typing: the caller's own record class (e.g. the reference's Instances) comes back
        gt_prop = rec(prop_i.image_size)
        gt_prop.proposal_boxes = gt_i
        gt_prop.objectness_logits = gt_logit_value * torch.ones(len(gt_i), device=dev)
        out.append(rec.cat([prop_i, gt_prop]))
    return out
