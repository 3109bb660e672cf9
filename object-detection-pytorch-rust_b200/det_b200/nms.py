"""NMS -- drop-in for ``batched_nms`` (reference python/src/utils.py:96-119) and the torchvision ``nms`` it wraps,
plus the batch-of-images form the kernels are built for.  Arithmetic: det_nms_batched (csrc/nms*.cu*)."""
from typing import Optional, Tuple

import torch

from . import _native as N

MODE_AUTO, MODE_PER_CATEGORY, MODE_OFFSET_TRICK = 0, 1, 2
MAX_BOXES_PER_IMAGE = 131071
MAX_CATEGORY = 32766


def nms_workspace_bytes(n: int, m: int) -> int:
    """Bytes of scratch det_nms_batched needs for n images of at most m boxes (det_nms_workspace_bytes)."""
    return int(N.fn("det_nms_workspace_bytes")(int(n), int(m)))


def nms_images(boxes: torch.Tensor, scores: torch.Tensor, categories: Optional[torch.Tensor],
               counts: Optional[torch.Tensor], iou_threshold: float, max_out: Optional[int] = None,
               mode: int = MODE_AUTO, workspace: Optional[torch.Tensor] = None,
               out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Whole-batch NMS.  boxes (N,M,4), scores (N,M), categories (N,M) int64 or None, counts (N) int32 or None.
    Returns (keep (N,max_out) int64 padded, keep_counts (N) int32); no host synchronisation.
    workspace (uint8, >= nms_workspace_bytes(N, M)) and out = (keep, keep_counts) let a serving loop run without any
    per-call allocation."""
    N.require_cuda(boxes, scores, categories, counts)
    assert boxes.dim() == 3 and boxes.shape[-1] == 4
    n, m = boxes.shape[0], boxes.shape[1]
    b, s = N.f32c(boxes), N.f32c(scores)
    c = None if categories is None else categories.detach().to(torch.int64).contiguous()
    k = None if counts is None else counts.detach().to(torch.int32).contiguous()
    max_out = m if max_out is None else int(max_out)
    if out is None:
        keep = torch.empty((n, max_out), dtype=torch.int64, device=b.device)
        keep_counts = torch.empty((n,), dtype=torch.int32, device=b.device)
    else:
        keep, keep_counts = out
        assert keep.shape == (n, max_out) and keep.dtype == torch.int64 and keep.is_contiguous()
        assert keep_counts.shape == (n,) and keep_counts.dtype == torch.int32
    if n == 0:
        return keep, keep_counts
    wsb = N.fn("det_nms_workspace_bytes")(n, m)
    ws = workspace
    if ws is None or ws.numel() < wsb or ws.dtype != torch.uint8 or ws.device != b.device:
        ws = torch.empty((wsb,), dtype=torch.uint8, device=b.device)
    with torch.cuda.device(b.device):
        N.call("det_nms_batched", N.ptr(b), N.ptr(s), N.ptr(c), N.ptr(k), n, m, float(iou_threshold), int(mode),
               max_out, N.ptr(keep), N.ptr(keep_counts), N.ptr(ws), wsb, N.stream())
    return keep, keep_counts


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """Same contract as the reference: int64 kept indices by descending score (python/src/utils.py:96)."""
    assert boxes.shape[-1] == 4
    if boxes.shape[0] == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    keep, cnt = nms_images(boxes[None], scores[None], idxs[None], None, iou_threshold)
    k = int(cnt.item())
    if k < 0:
        raise ValueError(f"batched_nms: category ids must lie in [0, {MAX_CATEGORY}]")
    return keep[0, :k]


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """torchvision.ops.nms contract (single category)."""
    if boxes.shape[0] == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    keep, cnt = nms_images(boxes[None], scores[None], None, None, iou_threshold, mode=MODE_PER_CATEGORY)
    return keep[0, :int(cnt.item())]
