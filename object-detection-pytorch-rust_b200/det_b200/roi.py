"""ROI-head side of target assignment -- drop-in for the matching / sampling half of the reference's ``ROIHeads``
(python/src/models/roi.py:15-193) and for ``subsample_labels`` (python/src/utils.py:34-76).

The IoU + Matcher arithmetic runs in the fused device kernels (det_match_anchors: no (G, P) matrix is materialised);
the fg/bg subsample keeps the reference's two ``torch.randperm`` calls so that, under the same torch seed, the sampled
indices are the reference's; the class-label convention is the reference's:
background = ``num_classes``, ignored = -1 (roi.py:84-95)."""
from typing import List, Optional, Tuple

import torch

from .matcher import Matcher
from .proposals import add_ground_truth_to_proposals
from .structures import Boxes, Instances


def subsample_labels(labels: torch.Tensor, num_samples: int, positive_fraction: float,
                     bg_label: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Indices of up to ``num_samples`` random elements of ``labels``: at most ``int(num_samples * positive_fraction)``
    positives (neither -1 nor ``bg_label``), the rest background -- same contract, counts and RNG consumption as
    reference python/src/utils.py:34-76."""
    positive = torch.nonzero((labels != -1) & (labels != bg_label), as_tuple=True)[0]
    negative = torch.nonzero(labels == bg_label, as_tuple=True)[0]
    num_pos = min(positive.numel(), int(num_samples * positive_fraction))
    num_neg = min(negative.numel(), num_samples - num_pos)
    pick_pos = torch.randperm(positive.numel(), device=positive.device)[:num_pos]
    pick_neg = torch.randperm(negative.numel(), device=negative.device)[:num_neg]
    return positive[pick_pos], negative[pick_neg]


class ROIHeads:
    """Matching / sampling logic of the reference's ROIHeads (roi.py:15-193); the per-region heads themselves are out
    of scope (the reference never wrote them)."""

    def __init__(self, num_classes: int, batch_size_per_image: int, positive_fraction: float,
                 proposal_matcher: Matcher, proposal_append_gt: bool = True):
        self.num_classes = int(num_classes)
        self.batch_size_per_image = int(batch_size_per_image)
        self.positive_fraction = float(positive_fraction)
        self.proposal_matcher = proposal_matcher
        self.proposal_append_gt = bool(proposal_append_gt)
        self.last_num_fg_samples: Optional[float] = None  # what the reference logs (roi.py:189-191)
        self.last_num_bg_samples: Optional[float] = None

    @classmethod
    def build(cls, conf):
        return cls(num_classes=conf.num_classes, batch_size_per_image=conf.batch_size_per_image,
                   positive_fraction=conf.positive_fraction, proposal_matcher=Matcher.build(conf.proposal_matcher),
                   proposal_append_gt=conf.proposal_append_gt)

    def label_proposals(self, proposal_boxes: torch.Tensor, gt_boxes: torch.Tensor,
                        gt_classes: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Deterministic half for one image: (matched gt index int64[P], matcher label int8[P], class label int64[P])
        with background = num_classes and ignored = -1 (roi.py:157-160 + :84-95)."""
        matched, labels, _, _ = self.proposal_matcher.match_boxes([gt_boxes.reshape(-1, 4)], proposal_boxes)
        matched, labels = matched[0], labels[0]
        if gt_classes.numel() > 0:
            cls = gt_classes.to(torch.int64)[matched]
            cls = torch.where(labels == 0, torch.full_like(cls, self.num_classes), cls)
            cls = torch.where(labels == -1, torch.full_like(cls, -1), cls)
        else:
            cls = torch.zeros_like(matched) + self.num_classes
        return matched, labels, cls

    def _sample_proposals(self, matched_idxs: torch.Tensor, matched_labels: torch.Tensor,
                          gt_classes: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """reference roi.py:68-105: class labels from the matching, then the fg/bg subsample."""
        if gt_classes.numel() > 0:
            cls = gt_classes.to(torch.int64)[matched_idxs]
            cls[matched_labels == 0] = self.num_classes
            cls[matched_labels == -1] = -1
        else:
            cls = torch.zeros_like(matched_idxs) + self.num_classes
        fg, bg = subsample_labels(cls, self.batch_size_per_image, self.positive_fraction, self.num_classes)
        sampled = torch.cat([fg, bg], dim=0)
        return sampled, cls[sampled]

    @torch.no_grad()
    def label_and_sample_proposals(self, proposals: List[Instances], targets: List[Instances]) -> List[Instances]:
        """reference roi.py:107-193: optional gt append, per image IoU -> Matcher -> sample, copy the gt_* fields of
        the matched targets onto the sampled proposals."""
        gt_boxes = [t.gt_boxes for t in targets]
        if self.proposal_append_gt:
            proposals = add_ground_truth_to_proposals(gt_boxes, proposals)
        out, num_fg, num_bg = [], [], []
        for props, tgt in zip(proposals, targets):
            has_gt = len(tgt) > 0
            gtb = tgt.gt_boxes.tensor if isinstance(tgt.gt_boxes, Boxes) else tgt.gt_boxes
            pb = props.proposal_boxes.tensor if isinstance(props.proposal_boxes, Boxes) else props.proposal_boxes
            matched, labels, _, _ = self.proposal_matcher.match_boxes([gtb.reshape(-1, 4)], pb)
            sampled, cls = self._sample_proposals(matched[0], labels[0], tgt.gt_classes)
            props = props[sampled]
            props.gt_classes = cls
            if has_gt:
                sampled_targets = matched[0][sampled]
                for name, value in tgt.get_fields().items():
                    if name.startswith("gt_") and not props.has(name):
                        props.set(name, value[sampled_targets])
            num_bg.append(int((cls == self.num_classes).sum()))
            num_fg.append(cls.numel() - num_bg[-1])
            out.append(props)
        if out:
            self.last_num_fg_samples = sum(num_fg) / len(num_fg)
            self.last_num_bg_samples = sum(num_bg) / len(num_bg)
        return out
