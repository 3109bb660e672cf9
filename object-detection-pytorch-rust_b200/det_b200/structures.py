"""Boundary types and pairwise overlap ops -- drop-in for the reference's ``python/src/structures``
(``Boxes``, ``Instances``, ``pairwise_iou``, ``pairwise_ioa``, ``pairwise_intersection``, ``matched_boxlist_iou``;
re-exported by ``python/src/structures/__init__.py:1-13``).  The overlap arithmetic runs in
``det_pairwise_overlap`` / ``det_matched_iou`` (csrc/box_ops.cu); the containers only hold tensors.
"""
from typing import Any, Dict, List, Tuple, Union

import torch

from . import _native as N


class Boxes:
    """Nx4 fp32 XYXY box container (reference: python/src/structures/boxes.py:4-170)."""

    def __init__(self, tensor: torch.Tensor):
        dev = tensor.device if isinstance(tensor, torch.Tensor) else torch.device("cpu")
        tensor = torch.as_tensor(tensor, dtype=torch.float32, device=dev)
        if tensor.numel() == 0:
            tensor = tensor.reshape((-1, 4)).to(dtype=torch.float32, device=dev)
        assert tensor.dim() == 2 and tensor.size(-1) == 4, tensor.size()
        self.tensor = tensor

    def clone(self) -> "Boxes":
        return Boxes(self.tensor.clone())

    def to(self, device) -> "Boxes":
        return Boxes(self.tensor.to(device=device))

    def area(self) -> torch.Tensor:
        t = self.tensor
        return (t[:, 2] - t[:, 0]) * (t[:, 3] - t[:, 1])

    def clip(self, box_size: Tuple[int, int]) -> None:
        """In-place clamp to [0,w]x[0,h]; asserts finiteness like boxes.py:60."""
        assert torch.isfinite(self.tensor).all(), "Box tensor contains infinite or NaN!"
        h, w = box_size
        self.tensor[:, 0::2].clamp_(min=0, max=w)
        self.tensor[:, 1::2].clamp_(min=0, max=h)

    def nonempty(self, threshold: float = 0.0) -> torch.Tensor:
        t = self.tensor
        return ((t[:, 2] - t[:, 0]) > threshold) & ((t[:, 3] - t[:, 1]) > threshold)

    def __getitem__(self, item) -> "Boxes":
        if isinstance(item, int):
            return Boxes(self.tensor[item].view(1, -1))
        b = self.tensor[item]
        assert b.dim() == 2, f"Indexing on Boxes with {item} failed to return a matrix!"
        return Boxes(b)

    def __len__(self) -> int:
        return self.tensor.shape[0]

    def __repr__(self) -> str:
        return f"Boxes({self.tensor})"

    def inside_box(self, box_size: Tuple[int, int], boundary_threshold: int = 0) -> torch.Tensor:
        h, w = box_size
        t = self.tensor
        return ((t[..., 0] >= -boundary_threshold) & (t[..., 1] >= -boundary_threshold)
                & (t[..., 2] < w + boundary_threshold) & (t[..., 3] < h + boundary_threshold))

    def get_centers(self) -> torch.Tensor:
        return (self.tensor[:, :2] + self.tensor[:, 2:]) / 2

    def scale(self, scale_x: float, scale_y: float) -> None:
        self.tensor[:, 0::2] *= scale_x
        self.tensor[:, 1::2] *= scale_y

    @classmethod
    def cat(cls, boxes_list: List["Boxes"]) -> "Boxes":
        assert isinstance(boxes_list, (list, tuple))
        if len(boxes_list) == 0:
            return cls(torch.empty(0))
        assert all(isinstance(b, Boxes) for b in boxes_list)
        return cls(torch.cat([b.tensor for b in boxes_list], dim=0))

    @property
    def device(self) -> torch.device:
        return self.tensor.device

    def __iter__(self):
        yield from self.tensor


def _as_tensor(b) -> torch.Tensor:
    return b.tensor if isinstance(b, Boxes) else b


def _pairwise(boxes1, boxes2, mode: int) -> torch.Tensor:
    b1, b2 = N.f32c(_as_tensor(boxes1)), N.f32c(_as_tensor(boxes2))
    N.require_cuda(b1, b2)
    n, m = b1.shape[0], b2.shape[0]
    out = torch.empty((n, m), dtype=torch.float32, device=b1.device)
    if n and m:
        with torch.cuda.device(b1.device):
            N.call("det_pairwise_overlap", N.ptr(b1), n, N.ptr(b2), m, mode, N.ptr(out), N.stream())
    return out


def pairwise_intersection(boxes1, boxes2) -> torch.Tensor:
    """[N,M] intersection areas (reference boxes.py:173)."""
    return _pairwise(boxes1, boxes2, 2)


def pairwise_iou(boxes1, boxes2) -> torch.Tensor:
    """[N,M] IoU, exactly 0 where the intersection is empty (reference boxes.py:193)."""
    return _pairwise(boxes1, boxes2, 0)


def pairwise_ioa(boxes1, boxes2) -> torch.Tensor:
    """[N,M] intersection over area(boxes2) (reference boxes.py:217)."""
    return _pairwise(boxes1, boxes2, 1)


def matched_boxlist_iou(boxes1, boxes2) -> torch.Tensor:
    """[N] IoU of matched pairs (reference boxes.py:235)."""
    assert len(boxes1) == len(boxes2), \
        "boxlists should have the samenumber of entries, got {}, {}".format(len(boxes1), len(boxes2))
    b1, b2 = N.f32c(_as_tensor(boxes1)), N.f32c(_as_tensor(boxes2))
    N.require_cuda(b1, b2)
    out = torch.empty((b1.shape[0],), dtype=torch.float32, device=b1.device)
    if b1.shape[0]:
        with torch.cuda.device(b1.device):
            N.call("det_matched_iou", N.ptr(b1), N.ptr(b2), b1.shape[0], N.ptr(out), N.stream())
    return out


def _cat_column(values: List[Any]) -> Any:
    first = values[0]
    if isinstance(first, torch.Tensor):
        return torch.cat(values, dim=0)
    if isinstance(first, list):
        return [x for v in values for x in v]
    if hasattr(type(first), "cat"):
        return type(first).cat(values)
    raise ValueError(f"Unsupported type {type(first)} for concatenation")


class Instances:
    """Per-image record: ``image_size`` plus any number of equally long columns (``proposal_boxes``,
    ``objectness_logits``, ``gt_boxes``, ``gt_classes`` ...).

    This is the boundary type of the reference's callables (python/src/structures/instances.py) and offers the calls
    the hot path and its callers use -- attribute access, ``has/get/set/remove/get_fields``, ``len``, indexing, ``to``,
    ``cat`` -- on an independent, much smaller implementation: columns are ordinary instance attributes.  The package
    itself only relies on duck typing (``image_size``, ``gt_boxes``, ``proposal_boxes``, ``objectness_logits``,
    ``gt_classes``, ``get_fields``), so the reference's own ``Instances`` objects can be passed in unchanged and
    functions that extend a caller's record return the caller's class."""

    def __init__(self, image_size: Tuple[int, int], **columns: Any):
        object.__setattr__(self, "image_size", image_size)
        for name, value in columns.items():
            setattr(self, name, value)

    def __setattr__(self, name: str, value: Any) -> None:
        assert name != "image_size", "image_size is fixed at construction"
        cols = self.get_fields()
        if cols and name not in cols:
            assert len(value) == len(self), f"Adding a field of length {len(value)} to a Instances of length {len(self)}"
        object.__setattr__(self, name, value)

    set = __setattr__

    def get_fields(self) -> Dict[str, Any]:
        return {k: v for k, v in vars(self).items() if k != "image_size"}

    def has(self, name: str) -> bool:
        return name in self.get_fields()

    def get(self, name: str) -> Any:
        return self.get_fields()[name]

    def remove(self, name: str) -> None:
        assert name != "image_size"
        object.__delattr__(self, name)

    def __len__(self) -> int:
        for value in self.get_fields().values():
            return len(value)
        raise NotImplementedError("Empty Instances does not support __len__!")

    def __iter__(self):
        raise NotImplementedError("`Instances` object is not iterable!")

    def _map(self, fn) -> "Instances":
        return type(self)(self.image_size, **{k: fn(v) for k, v in self.get_fields().items()})

    def to(self, *args: Any, **kwargs: Any) -> "Instances":
        return self._map(lambda v: v.to(*args, **kwargs) if hasattr(v, "to") else v)

    def __getitem__(self, item: Union[int, slice, torch.Tensor]) -> "Instances":
        if isinstance(item, int):
            n = len(self)
            if not -n <= item < n:
                raise IndexError("Instances index out of range!")
            item = slice(item % n, item % n + 1)
        return self._map(lambda v: v[item])

    @staticmethod
    def cat(instance_lists: List["Instances"]) -> "Instances":
        assert len(instance_lists) > 0
        first = instance_lists[0]
        if len(instance_lists) == 1:
            return first
        assert all(i.image_size == first.image_size for i in instance_lists)
        out = type(first)(first.image_size)
        for name in first.get_fields():
            out.set(name, _cat_column([i.get(name) for i in instance_lists]))
        return out

    def __repr__(self) -> str:
        cols = ", ".join(f"{k}: {v}" for k, v in self.get_fields().items())
        return (f"Instances(num_instances={len(self) if self.get_fields() else 0}, image_height={self.image_size[0]}, "
                f"image_width={self.image_size[1]}, fields=[{cols}])")
