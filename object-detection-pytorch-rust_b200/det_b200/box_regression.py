"""R-CNN box codec -- drop-in for ``Box2BoxTransform`` (reference python/src/models/components/box_regression.py:10-125).
Arithmetic: det_apply_deltas / det_get_deltas (csrc/box_ops.cu)."""
import math
from typing import Tuple

import torch

from . import _native as N

_DEFAULT_SCALE_CLAMP = math.log(1000.0 / 16)  # reference python/src/config/rpn.py:10


class Box2BoxTransform:
    def __init__(self, weights: Tuple[float, float, float, float] = (1.0, 1.0, 1.0, 1.0),
                 scale_clamp: float = _DEFAULT_SCALE_CLAMP):
        self.weights = tuple(float(w) for w in weights)
        self.scale_clamp = float(scale_clamp)

    @classmethod
    def build(cls, conf):
        return cls(conf.weights, conf.scale_clamp)

    def get_deltas(self, src_boxes: torch.Tensor, target_boxes: torch.Tensor) -> torch.Tensor:
        """(M,4),(M,4) -> (M,4) regression targets (dx,dy,dw,dh); asserts src widths > 0 like the reference."""
        assert isinstance(src_boxes, torch.Tensor), type(src_boxes)
        assert isinstance(target_boxes, torch.Tensor), type(target_boxes)
        N.require_cuda(src_boxes, target_boxes)
        s, t = N.f32c(src_boxes), N.f32c(target_boxes)
        out = torch.empty_like(s)
        flag = torch.zeros(1, dtype=torch.int32, device=s.device)
        if s.shape[0]:
            with torch.cuda.device(s.device):
                N.call("det_get_deltas", N.ptr(s), N.ptr(t), s.shape[0], *self.weights, N.ptr(out), N.ptr(flag),
                       N.stream())
        assert flag.item() == 0, "Input boxes to Box2BoxTransform are not valid!"
        return out

    def apply_deltas(self, deltas: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
        """(M,k*4),(M,4) -> (M,k*4) decoded boxes."""
        N.require_cuda(deltas, boxes)
        d, b = N.f32c(deltas), N.f32c(boxes)
        m = d.shape[0]
        k = (d.shape[1] // 4) if d.dim() == 2 else 1
        out = torch.empty_like(d)
        if m and k:
            with torch.cuda.device(d.device):
                N.call("det_apply_deltas", N.ptr(d), N.ptr(b), m, k, *self.weights, self.scale_clamp, N.ptr(out),
                       N.stream())
        return out
