"""FPN level assignment + ROIAlign -- drop-in for ``assign_boxes_to_levels``, ``convert_boxes_to_pooler_format``,
``ROIAlign`` and ``ROIPooler`` of the reference (python/src/models/modules/roi_poolers.py:15-331), forward and
backward (gradients flow to the feature maps, as through torchvision's roi_align).

The reference loops over the pyramid levels with nonzero / gather / torchvision ``roi_align`` / ``index_put_``; here
one kernel computes every box's level (det_roi_levels) and one kernel samples all levels (det_roi_align_levels).
Third-party arithmetic: torchvision.ops.roi_align (un-vendored) -- the parity oracle is the installed CPU kernel."""
import ctypes
import math
from typing import List, Sequence, Tuple, Union

import torch

from . import _native as N
from .structures import Boxes


def _box_tensors(box_lists: Sequence) -> List[torch.Tensor]:
    return [b.tensor if isinstance(b, Boxes) else b for b in box_lists]


def convert_boxes_to_pooler_format(box_lists: Sequence) -> torch.Tensor:
    """(M,5) rows (batch index, x0, y0, x1, y1) -- reference roi_poolers.py:141-166."""
    ts = _box_tensors(box_lists)
    return torch.cat([torch.cat((torch.full_like(t[:, :1], i), t), dim=1) for i, t in enumerate(ts)], dim=0)


def assign_boxes_to_levels(box_lists: Sequence, min_level: int, max_level: int, canonical_box_size: int,
                           canonical_level: int) -> torch.Tensor:
    """int64 (M,) level offsets from ``min_level`` -- reference roi_poolers.py:103-131 (Eqn.(1) of the FPN paper)."""
    boxes = N.f32c(torch.cat(_box_tensors(box_lists), dim=0))
    N.require_cuda(boxes)
    out = torch.empty((boxes.shape[0],), dtype=torch.int64, device=boxes.device)
    if boxes.shape[0]:
        with torch.cuda.device(boxes.device):
            N.call("det_roi_levels", N.ptr(boxes), boxes.shape[0], int(min_level), int(max_level),
                   float(canonical_box_size), int(canonical_level), N.ptr(out), N.stream())
    return out


def _level_table(tensors, scales):
    lv = (N.FeatureLevel * len(tensors))()
    for i, (f, s) in enumerate(zip(tensors, scales)):
        lv[i].data, lv[i].h, lv[i].w, lv[i].spatial_scale, lv[i].reserved = f.data_ptr(), f.shape[2], f.shape[3], float(s), 0
    return lv


class _RoiAlignLevels(torch.autograd.Function):
    """forward: det_roi_align_levels; backward: det_roi_align_levels_backward (gradients w.r.t. the feature maps)."""

    @staticmethod
    def forward(ctx, boxes, batch_index, level, scales, output_size, sampling_ratio, aligned, *feats):
        n, c = feats[0].shape[0], feats[0].shape[1]
        m = boxes.shape[0]
        out = torch.empty((m, c, output_size[0], output_size[1]), dtype=torch.float32, device=boxes.device)
        if m:
            lv = _level_table(feats, scales)
            with torch.cuda.device(boxes.device):
                N.call("det_roi_align_levels", ctypes.cast(lv, ctypes.c_void_p), len(feats), n, c, N.ptr(boxes),
                       N.ptr(batch_index), N.ptr(level), m, int(output_size[0]), int(output_size[1]), int(sampling_ratio),
                       int(bool(aligned)), N.ptr(out), N.stream())
        ctx.save_for_backward(boxes, batch_index, level if level is not None else torch.empty(0, device=boxes.device))
        ctx.meta = (scales, output_size, sampling_ratio, aligned, [tuple(f.shape) for f in feats], level is not None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        boxes, batch_index, level = ctx.saved_tensors
        scales, output_size, sampling_ratio, aligned, shapes, has_level = ctx.meta
        grads = [torch.zeros(sh, dtype=torch.float32, device=boxes.device) for sh in shapes]
        m = boxes.shape[0]
        if m:
            go = grad_out.contiguous().float()
            lv = _level_table(grads, scales)
            with torch.cuda.device(boxes.device):
                N.call("det_roi_align_levels_backward", ctypes.cast(lv, ctypes.c_void_p), len(grads), shapes[0][0],
                       shapes[0][1], N.ptr(boxes), N.ptr(batch_index), N.ptr(level if has_level else None), m,
                       int(output_size[0]), int(output_size[1]), int(sampling_ratio), int(bool(aligned)), N.ptr(go),
                       N.stream())
        return (None, None, None, None, None, None, None) + tuple(grads)


def _roi_align_levels(features: List[torch.Tensor], scales: Sequence[float], boxes: torch.Tensor,
                      batch_index: torch.Tensor, level, output_size: Tuple[int, int], sampling_ratio: int,
                      aligned: bool) -> torch.Tensor:
    N.require_cuda(boxes, *features)
    feats = [f.to(torch.float32).contiguous() for f in features]  # keeps the autograd graph of the feature maps
    n, c = feats[0].shape[0], feats[0].shape[1]
    for f in feats:
        assert f.dim() == 4 and f.shape[0] == n and f.shape[1] == c, "feature maps must share batch and channels"
    b = N.f32c(boxes)
    bi = batch_index.to(torch.int32).contiguous()
    out = _RoiAlignLevels.apply(b, bi, level, tuple(float(s) for s in scales), tuple(output_size), int(sampling_ratio),
                                bool(aligned), *feats)
    # the kernels compute in fp32; the result carries the feature maps' dtype like torchvision.ops.roi_align's does
    return out if features[0].dtype == torch.float32 else out.to(features[0].dtype)


class ROIAlign:
    """reference roi_poolers.py:15-98: ``forward(input NCHW, rois (B,5))`` -> (B, C, oh, ow)."""

    def __init__(self, output_size: Tuple[int, int], spatial_scale: float, sampling_ratio: int, aligned: bool = True):
        self.output_size = (int(output_size[0]), int(output_size[1]))
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.aligned = bool(aligned)

    @classmethod
    def build(cls, conf):
        return cls(conf.output_size, conf.spatial_scale, getattr(conf, "sampling_ration", getattr(conf, "sampling_ratio", 0)),
                   conf.aligned)

    def forward(self, input: torch.Tensor, rois: torch.Tensor) -> torch.Tensor:
        assert rois.dim() == 2 and rois.size(1) == 5
        r = rois.to(dtype=torch.float32)
        return _roi_align_levels([input], [self.spatial_scale], r[:, 1:5].contiguous(), r[:, 0], None, self.output_size,
                                 self.sampling_ratio, self.aligned)

    __call__ = forward

    def __repr__(self):
        return (f"ROIAlign(output_size={self.output_size}, spatial_scale={self.spatial_scale}, "
                f"sampling_ratio={self.sampling_ratio}, aligned={self.aligned})")


class ROIPooler:
    """reference roi_poolers.py:168-331: pools every box from the pyramid level its size selects."""

    def __init__(self, scales: List[float], sampling_ratio: int, output_size: Union[int, Tuple[int, int], List[int]],
                 type: str, canonical_box_size: int = 224, canonical_level: int = 4):
        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        assert len(output_size) == 2 and isinstance(output_size[0], int) and isinstance(output_size[1], int)
        self.output_size = (output_size[0], output_size[1])
        if type == "ROIAlign":
            self.aligned = False
        elif type == "ROIAlignV2":
            self.aligned = True
        elif type == "ROIPool":
            raise NotImplementedError("ROIPool (max pooling) is not part of this build; use ROIAlign / ROIAlignV2")
        else:
            raise ValueError("Unknown pooler type: {}".format(type))
        self.scales = [float(s) for s in scales]
        self.sampling_ratio = int(sampling_ratio)
        min_level, max_level = -(math.log2(scales[0])), -(math.log2(scales[-1]))
        assert math.isclose(min_level, int(min_level)) and math.isclose(max_level, int(max_level)), \
            "Featuremap stride is not power of 2!"
        self.min_level, self.max_level = int(min_level), int(max_level)
        assert len(scales) == self.max_level - self.min_level + 1, \
            "[ROIPooler] Sizes of input featuremaps do not form a pyramid!"
        assert 0 <= self.min_level <= self.max_level
        assert canonical_box_size > 0
        self.canonical_level, self.canonical_box_size = canonical_level, canonical_box_size

    @classmethod
    def build(cls, conf, scales: List[float]):
        return cls(scales=scales, output_size=conf.output_size, sampling_ratio=conf.sampling_ratio, type=conf.type,
                   canonical_box_size=conf.canonical_box_size, canonical_level=conf.canonical_level)

    def forward(self, x: List[torch.Tensor], box_lists: List[Boxes]) -> torch.Tensor:
        nl = len(self.scales)
        assert isinstance(x, list) and isinstance(box_lists, list), "Arguments to pooler must be lists"
        assert len(x) == nl, f"unequal value, num_level_assignments={nl}, but x is list of {len(x)} Tensors"
        assert len(box_lists) == x[0].size(0), \
            f"unequal value, x[0] batch dim 0 is {x[0].size(0)}, but box_list has length {len(box_lists)}"
        if len(box_lists) == 0:
            return torch.zeros((0, x[0].shape[1]) + self.output_size, device=x[0].device, dtype=x[0].dtype)
        ts = _box_tensors(box_lists)
        boxes = N.f32c(torch.cat(ts, dim=0))
        batch_index = torch.cat([torch.full((t.shape[0],), i, dtype=torch.int32, device=boxes.device)
                                 for i, t in enumerate(ts)])
        level = None
        if nl > 1:
            level = assign_boxes_to_levels(ts, self.min_level, self.max_level, self.canonical_box_size,
                                           self.canonical_level)
        return _roi_align_levels(x, self.scales, boxes, batch_index, level, self.output_size, self.sampling_ratio,
                                 self.aligned)

    __call__ = forward
