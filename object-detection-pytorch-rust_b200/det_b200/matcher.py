"""IoU -> (matched gt index, label) assignment -- drop-in for ``Matcher``
(reference python/src/models/components/matcher.py:8-120).  Arithmetic: det_match_quality (materialised matrix) and
det_match_anchors (fused IoU + matching for a whole batch, csrc/assign_loss.cu)."""
import ctypes
from typing import List, Optional, Sequence, Tuple

import torch

from . import _native as N


def _rule_arrays(thresholds: Sequence[float], labels: Sequence[int]):
    thr = (ctypes.c_float * max(len(thresholds), 1))(*[float(t) for t in thresholds])
    lab = (ctypes.c_int32 * len(labels))(*[int(l) for l in labels])
    return thr, lab


class Matcher:
    def __init__(self, thresholds: List[float], labels: List[int], allow_low_quality_matches: bool = True):
        thresholds = list(thresholds)
        assert thresholds[0] > 0
        bounds = [-float("inf")] + thresholds + [float("inf")]
        assert all(lo <= hi for lo, hi in zip(bounds[:-1], bounds[1:]))
        assert all(l in [-1, 0, 1] for l in labels)
        assert len(labels) == len(bounds) - 1
        self.thresholds = bounds            # same attribute layout as the reference (with the +-inf sentinels)
        self.labels = list(labels)
        self.allow_low_quality_matches = allow_low_quality_matches
        self._inner = thresholds            # finite thresholds handed to the kernels

    @classmethod
    def build(cls, conf):
        return cls(thresholds=conf.thresholds, labels=conf.labels,
                   allow_low_quality_matches=conf.allow_low_quality_matches)

    def __call__(self, match_quality_matrix: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(G,R) quality matrix -> (int64[R] matched gt index, int8[R] label in {-1,0,1})."""
        q = match_quality_matrix
        assert q.dim() == 2
        N.require_cuda(q)
        q = N.f32c(q)
        g, r = q.shape
        matches = torch.empty((r,), dtype=torch.int64, device=q.device)
        labels = torch.empty((r,), dtype=torch.int8, device=q.device)
        if r == 0:
            return matches, labels
        flag = torch.zeros(1, dtype=torch.int32, device=q.device)
        ws = torch.empty((max(g, 1),), dtype=torch.float32, device=q.device)
        thr, lab = _rule_arrays(self._inner, self.labels)
        with torch.cuda.device(q.device):
            N.call("det_match_quality", N.ptr(q), g, r, thr, lab, len(self._inner),
                   int(self.allow_low_quality_matches), N.ptr(matches), N.ptr(labels), N.ptr(flag), N.ptr(ws),
                   ws.numel() * 4, N.stream())
        assert flag.item() == 0  # reference: assert torch.all(match_quality_matrix >= 0)
        return matches, labels

    def match_boxes(self, gt_boxes: Sequence[torch.Tensor], anchors: torch.Tensor, return_iou: bool = False, grid=None,
                    with_stats: bool = False):
        """Fused batch form: per image IoU(gt_i, anchors) + matching without materialising the matrices.
        gt_boxes: list of (G_i,4) tensors; anchors (R,4).  Returns matched (N,R) int64, labels (N,R) int8
        [, matched IoU (N,R)] [, MatchStats], and the packed gt table (sum_G,4) + int32 offsets (N+1) used by the loss
        kernel.  grid / with_stats: see match_packed."""
        N.require_cuda(anchors, *gt_boxes)
        a = N.f32c(anchors)
        table, offsets = self.pack_gt(gt_boxes, anchors.device)
        return self.match_packed(table, offsets, len(gt_boxes), a, return_iou, grid, with_stats) + (table, offsets)

    @staticmethod
    def pack_gt(gt_boxes: Sequence[torch.Tensor], device):
        """list of (G_i,4) tensors -> packed (sum_G,4) fp32 table + int32 offsets (N+1) on `device`."""
        n = len(gt_boxes)
        table = (N.f32c(torch.cat([g.reshape(-1, 4) for g in gt_boxes], dim=0)) if n
                 else torch.zeros((0, 4), dtype=torch.float32, device=device))
        off_host = [0]
        for g in gt_boxes:
            off_host.append(off_host[-1] + int(g.shape[0]))
        offsets = torch.tensor(off_host, dtype=torch.int32).to(device, non_blocking=True)
        return table, offsets

    def match_packed(self, table: torch.Tensor, offsets: torch.Tensor, n: int, anchors: torch.Tensor,
                     return_iou: bool = False, grid=None, with_stats: bool = False):
        """As match_boxes, for an already packed gt table (sum_G,4) + device int32 offsets (n+1).

        grid = (levels, a) with levels = [(h, w, stride), ...] declares that `anchors` is the (level, h, w, a) table of
        AnchorGenerator (anchor_generators.py:158-179): the one-pass grid kernels (det_match_grid) are used -- same
        results bit for bit.  with_stats (grid only) additionally returns a MatchStats (per-image positive / ignored
        counts and the list of positives) that subsample_labels_ turns into an O(samples) subsample."""
        dev = anchors.device
        r, sum_g = anchors.shape[0], table.shape[0]
        matched = torch.empty((n, r), dtype=torch.int64, device=dev)
        labels = torch.empty((n, r), dtype=torch.int8, device=dev)
        miou = torch.empty((n, r), dtype=torch.float32, device=dev) if return_iou else None
        stats = None
        if n and r:
            thr, lab = _rule_arrays(self._inner, self.labels)
            if grid is not None and grid_supported(grid, r):
                levels, a = grid
                lv = (N.AnchorLevel * len(levels))()
                row = 0
                for i, (h, w, stride) in enumerate(levels):
                    lv[i].h, lv[i].w, lv[i].stride, lv[i].reserved, lv[i].first_row = int(h), int(w), int(stride), 0, row
                    row += int(h) * int(w) * a
                wsb = N.fn("det_match_grid_workspace_bytes")(n, sum_g, ctypes.cast(lv, ctypes.c_void_p), len(levels))
                ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
                if with_stats and r < (1 << 24):
                    stats = MatchStats(torch.empty((n, 4), dtype=torch.int32, device=dev),
                                       torch.empty((n, MatchStats.LIST_CAP), dtype=torch.int32, device=dev))
                with torch.cuda.device(dev):
                    N.call("det_match_grid", N.ptr(table), N.ptr(offsets), n, sum_g, N.ptr(anchors), r,
                           ctypes.cast(lv, ctypes.c_void_p), len(levels), int(a), thr, lab, len(self._inner),
                           int(self.allow_low_quality_matches), N.ptr(matched), N.ptr(labels), N.ptr(miou),
                           N.ptr(stats.counts if stats else None), N.ptr(stats.pos_list if stats else None),
                           MatchStats.LIST_CAP, N.ptr(ws), wsb, N.stream())
            else:
                wsb = N.fn("det_match_workspace_bytes")(n, r, sum_g)
                ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
                with torch.cuda.device(dev):
                    N.call("det_match_anchors", N.ptr(table), N.ptr(offsets), n, sum_g, N.ptr(anchors), r, thr, lab,
                           len(self._inner), int(self.allow_low_quality_matches), N.ptr(matched), N.ptr(labels),
                           N.ptr(miou), N.ptr(ws), wsb, N.stream())
        out = (matched, labels, miou) if return_iou else (matched, labels)
        return out + (stats,) if with_stats else out


class MatchStats:
    """Per-image by-products of det_match_grid: counts (n,4) int32 = {#positives, #ignored, 0, 0} and pos_list
    (n, LIST_CAP) int32 (anchor row | label << 24).  Valid for the labels the same call wrote, until they are changed."""
    LIST_CAP = 1024

    def __init__(self, counts: torch.Tensor, pos_list: torch.Tensor):
        self.counts, self.pos_list = counts, pos_list


def grid_supported(grid, r: int) -> bool:
    """det_match_grid's documented limits (otherwise the generic kernels run)."""
    levels, a = grid
    return (a in (1, 3, 9) and 1 <= len(levels) <= 8 and len(levels) * a <= 32
            and sum(int(h) * int(w) * a for h, w, _ in levels) == r)


def subsample_labels_(labels: torch.Tensor, num_samples: int, positive_fraction: float, seed: int,
                      stats: Optional[MatchStats] = None, return_samples: bool = False):
    """In-place device fg/bg subsample of (N,R) int8 labels in {-1,0,1}: keeps min(#pos, int(S*f)) positives and
    min(#neg, S-#pos) negatives per image, chosen uniformly at random (counter-based hash of seed,image,anchor);
    everything else becomes -1.  Same counts/distribution as reference subsample_labels + _subsample_labels
    (python/src/utils.py:34, models/rpn.py:108); the random stream necessarily differs from torch.randperm.

    stats (from Matcher.match_packed(..., with_stats=True) on these very labels): the O(samples) form
    det_subsample_labels_grid -- same result bit for bit.  return_samples (needs stats): also returns the per-image
    sample lists (N, num_samples) int32 (anchor row | label << 24) and counts (N) int32 for the sampled loss."""
    N.require_cuda(labels)
    assert labels.dtype == torch.int8 and labels.is_contiguous() and labels.dim() == 2
    n, r = labels.shape
    samples = counts = None
    if stats is not None and n and r:
        cap = max(int(num_samples), 1)
        if return_samples and r < (1 << 24):
            samples = torch.empty((n, cap), dtype=torch.int32, device=labels.device)
            counts = torch.empty((n,), dtype=torch.int32, device=labels.device)
        with torch.cuda.device(labels.device):
            N.call("det_subsample_labels_grid", N.ptr(labels), n, r, int(num_samples), float(positive_fraction),
                   int(seed) & 0xFFFFFFFFFFFFFFFF, N.ptr(stats.counts), N.ptr(stats.pos_list), MatchStats.LIST_CAP,
                   N.ptr(samples), N.ptr(counts), cap, N.stream())
    elif n and r:
        with torch.cuda.device(labels.device):
            N.call("det_subsample_labels", N.ptr(labels), n, r, int(num_samples), float(positive_fraction),
                   int(seed) & 0xFFFFFFFFFFFFFFFF, N.stream())
    return (labels, samples, counts) if return_samples else labels
