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


def assert_boxes_close(got, want, rtol=1e-5):
    """fp32 boxes within `rtol` of the box's own scale (largest |coordinate| or side): corners are differences of a
    centre and a half-size, so the error of the transcendental shows up relative to those, not to the corner."""
    assert got.shape == want.shape
    g, w = got.reshape(-1, 4).double(), want.reshape(-1, 4).double()
    side = torch.maximum((w[:, 2] - w[:, 0]).abs(), (w[:, 3] - w[:, 1]).abs())
    scale = torch.maximum(w.abs().amax(dim=1), side).clamp_min(1.0)
    err = (g - w).abs().amax(dim=1)
    bad = err > rtol * scale
    assert not bool(bad.any()), f"{int(bad.sum())} boxes off by more than {rtol} rel; worst {float((err / scale).max()):.3e}"
