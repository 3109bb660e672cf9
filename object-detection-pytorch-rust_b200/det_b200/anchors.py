"""Grid anchor generation -- drop-in for ``AnchorGenerator`` (reference
python/src/models/modules/anchor_generators.py:84-238).  Anchors of a level are produced by det_grid_anchors;
the decode kernels never read them (they synthesise anchors from indices), this class exists for API parity and
for the assignment kernel, which takes the (R,4) anchor table."""
import math
from typing import List, Sequence, Tuple, Union

import torch

from . import _native as N
from .structures import Boxes


def _broadcast_params(params, num_features: int, name: str):
    assert isinstance(params, Sequence), f"{name} in anchor generator has to be a list! Got {params}."
    assert len(params), f"{name} in anchor generator cannot be empty!"
    if not isinstance(params[0], Sequence):
        return [params] * num_features
    if len(params) == 1:
        return list(params) * num_features
    assert len(params) == num_features, (
        f"Got {name} of length {len(params)} in anchor generator, but the number of input features is {num_features}!")
    return params


def generate_cell_anchors(sizes=(32, 64, 128, 256, 512), aspect_ratios=(0.5, 1, 2)) -> torch.Tensor:
    """Anchors centred on (0,0), python double math (anchor_generators.py:181-210)."""
    rows = []
    for size in sizes:
        area = size ** 2.0
        for r in aspect_ratios:
            w = math.sqrt(area / r)
            h = r * w
            rows.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
    return torch.tensor(rows)


class AnchorGenerator:
    box_dim = 4

    def __init__(self, strides: List[int], sizes=((32,), (64,), (128,), (256,), (512,)),
                 aspect_ratios=((0.5, 1.0, 2.0),), offset: float = 0.0, box_dim: int = 4):
        self.strides = list(strides)
        self.num_features = len(self.strides)
        sizes = _broadcast_params(sizes, self.num_features, "sizes")
        aspect_ratios = _broadcast_params(aspect_ratios, self.num_features, "aspect_ratios")
        self.cell_anchors = [generate_cell_anchors(s, a).float() for s, a in zip(sizes, aspect_ratios)]
        self.offset = offset
        assert 0.0 <= self.offset < 1.0, self.offset
        assert box_dim == 4
        self._dev_cells = {}

    @classmethod
    def build(cls, conf, strides: List[int]):
        return cls(strides=strides, sizes=conf.sizes, aspect_ratios=conf.aspect_ratios, offset=conf.offset,
                   box_dim=conf.box_dim)

    @property
    def num_anchors(self) -> List[int]:
        return [len(c) for c in self.cell_anchors]

    num_cell_anchors = num_anchors

    def device_cell_anchors(self, device) -> List[torch.Tensor]:
        key = str(device)
        if key not in self._dev_cells:
            self._dev_cells[key] = [c.to(device).contiguous() for c in self.cell_anchors]
        return self._dev_cells[key]

    def grid_anchors(self, grid_sizes: List[Tuple[int, int]], device) -> List[torch.Tensor]:
        cells = self.device_cell_anchors(device)
        out = []
        with torch.cuda.device(device):
            for (h, w), stride, cell in zip(grid_sizes, self.strides, cells):
                a = cell.shape[0]
                t = torch.empty((h * w * a, 4), dtype=torch.float32, device=device)
                if t.numel():
                    N.call("det_grid_anchors", N.ptr(cell), a, int(h), int(w), int(stride), float(self.offset),
                           N.ptr(t), N.stream())
                t._det_grid = (int(h), int(w), int(stride), int(a))  # layout tag read by the grid matcher (never guessed)
                out.append(t)
        return out

    def grid_layout(self, grid_sizes: List[Tuple[int, int]]):
        """(levels, a) for Matcher.match_packed(grid=...), or None when the levels differ in anchors per position."""
        na = self.num_anchors
        if len(set(na)) != 1:
            return None
        return [(int(h), int(w), int(s)) for (h, w), s in zip(grid_sizes, self.strides)], na[0]

    def forward(self, features: List[torch.Tensor]) -> List[Boxes]:
        N.require_cuda(*features)
        sizes = [tuple(f.shape[-2:]) for f in features]
        return [Boxes(t) for t in self.grid_anchors(sizes, features[0].device)]

    __call__ = forward
