"""CPU: the C restatement of greedy NMS (oracle/nms_ref.c) pinned against the installed torchvision CPU kernel
(third-party arithmetic the reference reaches through python/src/utils.py:110,115)."""
import pytest
import torch

from oracle import ref_torch as O
from tests.util import gen, rand_boxes

tv = pytest.importorskip("torchvision")


@pytest.mark.parametrize("seed", range(40))
def test_nms_matches_torchvision(seed):
    g = gen(seed)
    n = int(torch.randint(1, 600, (1,), generator=g))
    b = rand_boxes(n, 300.0, g)
    if seed % 3 == 0:
        b = b.round()
    s = torch.rand(n, generator=g)
    if seed % 4 == 0:
        s = (s * 8).round() / 8  # heavy ties: torchvision sorts stably
    thr = [0.5, 0.7, 0.3, 0.05][seed % 4]
    assert torch.equal(O.nms(b, s, thr), tv.ops.nms(b, s, thr))


@pytest.mark.parametrize("seed", range(20))
def test_nms_nan_inf_matches_torchvision(seed):
    g = gen(1000 + seed)
    n = 60
    b, s = rand_boxes(n, 100.0, g), torch.rand(n, generator=g)
    b[torch.rand(n, 4, generator=g) < 0.05] = float("nan")
    b[torch.rand(n, 4, generator=g) < 0.03] = float("inf")
    s[torch.rand(n, generator=g) < 0.1] = float("nan")
    assert torch.equal(O.nms(b, s, 0.5), tv.ops.nms(b, s, 0.5))


def test_threshold_semantics():
    s = torch.tensor([1.0, 0.5])
    b = torch.tensor([[0, 0, 2, 1], [0, 0, 1, 1.0]])  # IoU exactly 0.5: kept (strict >)
    assert O.nms(b, s, 0.5).tolist() == tv.ops.nms(b, s, 0.5).tolist() == [0, 1]
    b = torch.tensor([[0, 0, 10, 1], [0, 0, 7, 1.0]])  # IoU == float32(0.7) < double 0.7: kept
    assert O.nms(b, s, 0.7).tolist() == tv.ops.nms(b, s, 0.7).tolist() == [0, 1]
    assert O.nms(b, s, 0.6999).tolist() == tv.ops.nms(b, s, 0.6999).tolist() == [0]
    # threshold strictly between float32(0.7) and its predecessor
    thr = float(torch.tensor(0.7).item()) - 1e-9
    assert O.nms(b, s, thr).tolist() == tv.ops.nms(b, s, thr).tolist() == [0]


def test_empty():
    assert O.nms(torch.zeros(0, 4), torch.zeros(0), 0.5).numel() == 0
    assert O.batched_nms(torch.zeros(0, 4), torch.zeros(0), torch.zeros(0, dtype=torch.int64), 0.5).numel() == 0
