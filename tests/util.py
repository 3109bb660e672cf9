"""Seeded synthetic inputs shared by the tests (CPU tensors; GPU tests copy them so both sides see the same bits)."""
import torch


def gen(seed):
    return torch.Generator().manual_seed(seed)


def rand_boxes(n, frame=640.0, g=None, wh_frac=0.2):
    """XYXY boxes: xy ~ U(0, 0.8*frame), wh ~ U(1, wh_frac*frame + 1)  (SURVEY.md 8d Cfg3 recipe)."""
    xy = torch.rand(n, 2, generator=g) * frame * 0.8
    wh = torch.rand(n, 2, generator=g) * frame * wh_frac + 1
    return torch.cat([xy, xy + wh], 1)


def distinct_scores(n, g=None):
    """fp32 scores without ties (the reference's unstable sorts are only defined up to tie order)."""
    return (torch.randperm(n, generator=g).float() + 0.5) / max(n, 1)


def ulp_diff(a, b):
    ai = a.contiguous().view(torch.int32).long()
    bi = b.contiguous().view(torch.int32).long()
    ai = torch.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = torch.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return (ai - bi).abs().max().item() if a.numel() else 0
