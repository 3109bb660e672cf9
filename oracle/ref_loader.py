"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* upstream reference.

Imports the reference's own Python functions from ``/root/reference`` so that
(a) the restatement in ``oracle/ref_torch.py`` can be validated against them and
(b) golden vectors under ``tests/golden/`` can be generated (``tests/golden/make_golden.py``).

``/root/reference`` only exists in the build container, never on the GPU box, so
nothing in the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may import this module.

The reference does not import as-is:
  * ``fvcore`` is absent -> a minimal stand-in for the three symbols the reference
    pulls from it (``python/src/models/components/box_regression.py:4``,
    ``python/src/structures/storage.py:6``) is installed into ``sys.modules``.
    The stand-in follows fvcore's published definitions (smooth-L1 with the
    beta<1e-5 => L1 switch; GIoU with eps=1e-7); it is "from memory", the default
    RPN configuration (beta=0 => pure L1) does not depend on any ambiguity.
  * ``Logs.get_instance`` crashes (``storage.py:37``), so it is replaced by a no-op
    sink; the arithmetic of ``losses`` is untouched.
"""
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("DET_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "python", "src"))


def _install_fvcore_stub():
    if "fvcore" in sys.modules:
        return
    fv = types.ModuleType("fvcore")
    fv_nn = types.ModuleType("fvcore.nn")
    fv_common = types.ModuleType("fvcore.common")
    fv_hist = types.ModuleType("fvcore.common.history_buffer")

    def smooth_l1_loss(input, target, beta, reduction="none"):
        if beta < 1e-5:
            loss = torch.abs(input - target)
        else:
            n = torch.abs(input - target)
            loss = torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta)
        if reduction == "mean":
            loss = loss.mean() if loss.numel() > 0 else 0.0 * loss.sum()
        elif reduction == "sum":
            loss = loss.sum()
        return loss

    def giou_loss(boxes1, boxes2, reduction="none", eps=1e-7):
        x1, y1, x2, y2 = boxes1.unbind(dim=-1)
        x1g, y1g, x2g, y2g = boxes2.unbind(dim=-1)
        xkis1 = torch.max(x1, x1g)
        ykis1 = torch.max(y1, y1g)
        xkis2 = torch.min(x2, x2g)
        ykis2 = torch.min(y2, y2g)
        intsctk = torch.zeros_like(x1)
        mask = (ykis2 > ykis1) & (xkis2 > xkis1)
        intsctk[mask] = (xkis2[mask] - xkis1[mask]) * (ykis2[mask] - ykis1[mask])
        unionk = (x2 - x1) * (y2 - y1) + (x2g - x1g) * (y2g - y1g) - intsctk
        iouk = intsctk / (unionk + eps)
        xc1 = torch.min(x1, x1g)
        yc1 = torch.min(y1, y1g)
        xc2 = torch.max(x2, x2g)
        yc2 = torch.max(y2, y2g)
        area_c = (xc2 - xc1) * (yc2 - yc1)
        miouk = iouk - ((area_c - unionk) / (area_c + eps))
        loss = 1 - miouk
        if reduction == "mean":
            loss = loss.mean() if loss.numel() > 0 else 0.0 * loss.sum()
        elif reduction == "sum":
            loss = loss.sum()
        return loss

    class HistoryBuffer:  # never exercised: Logs is replaced by a sink below
        def __init__(self, max_length=1000000):
            self._max_length = max_length

    fv_nn.smooth_l1_loss = smooth_l1_loss
    fv_nn.giou_loss = giou_loss
    fv_hist.HistoryBuffer = HistoryBuffer
    fv.nn = fv_nn
    fv.common = fv_common
    fv_common.history_buffer = fv_hist
    sys.modules.update({
        "fvcore": fv, "fvcore.nn": fv_nn,
        "fvcore.common": fv_common, "fvcore.common.history_buffer": fv_hist,
    })


class _Sink:
    def put_scalar(self, *a, **k):
        pass

    def put_scalars(self, *a, **k):
        pass


_LOADED = None


def load():
    """Return a namespace with the reference callables on the hot path."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    _install_fvcore_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from python.src.structures import boxes as r_boxes
    from python.src.structures import Instances, Logs
    from python.src.models.components import box_regression as r_breg
    from python.src.models.components import matcher as r_matcher
    from python.src.models import utils as r_mutils
    from python.src import utils as r_utils
    from python.src.models.modules import anchor_generators as r_anchor
    from python.src.models import rpn as r_rpn
    from python.src.models import roi as r_roi
    from python.src import config as r_config

    Logs.get_instance = staticmethod(lambda: _Sink())
    ns = types.SimpleNamespace(
        Boxes=r_boxes.Boxes, Instances=Instances,
        pairwise_iou=r_boxes.pairwise_iou, pairwise_ioa=r_boxes.pairwise_ioa,
        pairwise_intersection=r_boxes.pairwise_intersection,
        matched_boxlist_iou=r_boxes.matched_boxlist_iou,
        Box2BoxTransform=r_breg.Box2BoxTransform,
        dense_box_regression_loss=r_breg._dense_box_regression_loss,
        Matcher=r_matcher.Matcher,
        find_top_rpn_proposals=r_mutils.find_top_rpn_proposals,
        add_ground_truth_to_proposals=r_mutils.add_ground_truth_to_proposals,
        batched_nms=r_utils.batched_nms, subsample_labels=r_utils.subsample_labels,
        AnchorGenerator=r_anchor.AnchorGenerator,
        RegionProposalNetwork=r_rpn.RegionProposalNetwork,
        ROIHeads=r_roi.ROIHeads,
        config=r_config, utils=r_utils,
    )
    _LOADED = ns
    return ns
