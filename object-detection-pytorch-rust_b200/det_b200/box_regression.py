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
        """(M,k*4),(M,4) -> (M,k*4) decoded boxes.  Differentiable like the reference's (autograd through
        det_apply_deltas_backward) whenever deltas or boxes require grad."""
        N.require_cuda(deltas, boxes)
        if torch.is_grad_enabled() and (deltas.requires_grad or boxes.requires_grad):
            return _ApplyDeltas.apply(deltas, boxes, self.weights, self.scale_clamp)
        return _apply_deltas_forward(N.f32c(deltas), N.f32c(boxes), self.weights, self.scale_clamp)


def _apply_deltas_forward(d: torch.Tensor, b: torch.Tensor, weights, scale_clamp: float) -> torch.Tensor:
    m = d.shape[0]
    k = (d.shape[1] // 4) if d.dim() == 2 else 1
    out = torch.empty_like(d)
    if m and k:
        with torch.cuda.device(d.device):
            N.call("det_apply_deltas", N.ptr(d), N.ptr(b), m, k, *weights, scale_clamp, N.ptr(out), N.stream())
    return out


class _ApplyDeltas(torch.autograd.Function):
    @staticmethod
    def forward(ctx, deltas, boxes, weights, scale_clamp):
        d, b = N.f32c(deltas), N.f32c(boxes)
        ctx.weights, ctx.scale_clamp = weights, scale_clamp
        ctx.dtypes = (deltas.dtype, boxes.dtype)
        ctx.save_for_backward(d, b)
        return _apply_deltas_forward(d, b, weights, scale_clamp)

    @staticmethod
    def backward(ctx, grad_out):
        d, b = ctx.saved_tensors
        m = d.shape[0]
        k = (d.shape[1] // 4) if d.dim() == 2 else 1
        need_d, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gd = torch.zeros_like(d) if need_d else None
        gb = torch.zeros_like(b) if need_b else None
        if m and k:
            go = N.f32c(grad_out)
            with torch.cuda.device(d.device):
                N.call("det_apply_deltas_backward", N.ptr(d), N.ptr(b), N.ptr(go), m, k, *ctx.weights, ctx.scale_clamp,
                       N.ptr(gd), N.ptr(gb), N.stream())
        return (gd.to(ctx.dtypes[0]) if need_d else None, gb.to(ctx.dtypes[1]) if need_b else None, None, None)
