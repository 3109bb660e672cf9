"""CPU, world_size 2, gloo: the data-parallel host logic -- batch sharding, global normaliser, the single SUM
all-reduce of the loss/count vector.  Each rank evaluates its shard with the oracle (no GPU here) exactly the way
bench.py / RegionProposalNetwork.fused_losses combine the kernel's sums, and the result must equal the oracle on the
whole batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import gen, rand_boxes


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    from oracle import ref_torch as O
    g = gen(5)
    cells = [O.cell_anchors(s, [0.5, 1.0, 2.0]) for s in ([32], [64], [128])]
    anc = torch.cat(O.grid_anchors([(16, 16), (8, 8), (4, 4)], [8, 16, 32], cells, 0.0), 0)
    n = 6
    gts = [rand_boxes(3 + i, 128.0, g, 0.5).clamp(max=128.0) for i in range(n)]
    labs, idxs = O.label_anchors(anc, gts)
    mb = torch.stack([gt[i] for gt, i in zip(gts, idxs)])
    logits = torch.randn(n, anc.shape[0], generator=g)
    deltas = torch.randn(n, anc.shape[0], 4, generator=g) * 0.5
    return O, anc, torch.stack(labs), mb, logits, deltas


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import det_b200
    r, w, _ = det_b200.dist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    O, anc, labs, mb, logits, deltas = _problem()
    n = labs.shape[0]
    lo, hi = det_b200.dist.shard_range(n, rank, world)
    n_global = det_b200.dist.global_num_images(hi - lo)
    assert n_global == n
    part = O.rpn_losses(anc, logits[lo:hi], labs[lo:hi], deltas[lo:hi], mb[lo:hi], batch_size_per_image=256)
    # local sums scaled by the GLOBAL normaliser: loss_local * n_local / n_global
    sums = torch.zeros(8)
    sums[0] = part["cls_loss"] * (hi - lo) / n_global
    sums[1] = part["loc_loss"] * (hi - lo) / n_global
    sums[2], sums[3] = part["num_pos"], part["num_neg"]
    late = sums.clone()
    det_b200.dist.allreduce_sums_(sums)
    work = det_b200.dist.allreduce_sums_async(late)  # the one-step-late form used by the training loop
    work.wait()
    assert torch.equal(late, sums)
    # logging-cadence reducer: 3 steps accumulated, one all-reduce
    red = det_b200.dist.SumsReducer(every=3)
    local = torch.full((8,), float(rank + 1))
    assert red.add(local) is None and red.add(local) is None
    got = red.add(local)
    red.flush()
    assert got is not None and torch.equal(got, torch.full((8,), 3.0 * sum(range(1, world + 1))))
    if rank == 0:
        torch.save(sums, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_loss_allreduce_equals_full_batch(tmp_path):
    out = str(tmp_path / "sums.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    sums = torch.load(out)
    O, anc, labs, mb, logits, deltas = _problem()
    full = O.rpn_losses(anc, logits, labs, deltas, mb, batch_size_per_image=256)
    torch.testing.assert_close(sums[0], full["cls_loss"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(sums[1], full["loc_loss"], rtol=1e-5, atol=1e-7)
    assert int(sums[2]) == full["num_pos"] and int(sums[3]) == full["num_neg"]
