"""Race / hazard evidence without compute-sanitizer (the tool is closed on the GPU pool: profiles/r02/sanitizer_refused.txt).

Every kernel that synchronises through shared-memory atomics, grid-wide barriers, tickets ("last CTA arrives") or spin
waits is run many times on the SAME inputs -- alone, and with a second stream keeping the SMs busy so that CTA
scheduling and arrival order differ from run to run -- and every output must be bit-identical across runs.  A data race
in one of these kernels shows up as run-to-run differences (kept lists, counts, sums); the parity tests next to this
file pin the values themselves against the oracle.  Kernels covered: yolo_fast_kernel (smem atomicOr / atomicAdd),
large_huge_segments_kernel (cooperative grid barrier), dense_detect (per-image counters + NMS CTA), subsample kernels,
match_grid / assign_candidates / subsample_lazy (list appends through atomics: order-free results), the loss kernels'
last-CTA finalisation, peer_sums_exchange (stamps + spin)."""
import pytest
import torch

from tests.util import gen, rand_boxes

pytestmark = pytest.mark.gpu

REPEATS = 25


@pytest.fixture(scope="module")
def det():
    import det_b200
    return det_b200


class Noise:
    """A side stream that keeps a varying number of SMs busy while the kernel under test runs."""

    def __init__(self):
        self.stream = torch.cuda.Stream()
        self.a = torch.randn(1 << 22, device="cuda")
        self.k = 0

    def kick(self):
        self.k += 1
        with torch.cuda.stream(self.stream):
            for _ in range(1 + self.k % 3):
                self.a.mul_(1.0001).add_(0.5)


def run_repeated(fn, snapshot):
    """fn() launches; snapshot() -> list of tensors.  Bitwise comparison of every repeat with the first run."""
    noise = Noise()
    fn()
    torch.cuda.synchronize()
    first = [t.clone() for t in snapshot()]
    for rep in range(REPEATS):
        if rep % 2:
            noise.kick()
        fn()
        torch.cuda.synchronize()
        for a, b in zip(first, snapshot()):
            assert a.dtype == b.dtype and a.shape == b.shape
            assert torch.equal(a.contiguous().view(torch.uint8), b.contiguous().view(torch.uint8)), f"repeat {rep} differs"


def test_yolo_fast_kernel_is_deterministic(det):
    yh = det.YoloGridHead(7, 2, 20, (448, 448))
    head = torch.randn(64, 7, 7, 30, generator=gen(1)).cuda()
    out = yh.detect(head, 0.25, 0.5, max_det=300)
    run_repeated(lambda: yh.detect(head, 0.25, 0.5, max_det=300, out=out),
                 lambda: [out["flat"], out["count"], out["boxes"], out["scores"]])
    out2 = yh.detect(head, 0.05, 0.5, max_det=40)  # all three tiers
    run_repeated(lambda: yh.detect(head, 0.05, 0.5, max_det=40, out=out2), lambda: [out2["flat"], out2["count"]])


@pytest.mark.parametrize("m,ncat", [(6000, 1), (12000, 2), (5000, 80)])
def test_large_nms_paths_are_deterministic(det, m, ncat):
    g = gen(m)
    b = rand_boxes(m, 1024.0, g).cuda()
    s = torch.rand(m, generator=g).cuda()
    c = torch.randint(0, ncat, (m,), generator=g).cuda()
    res = {}

    def fn():
        res["keep"] = det.batched_nms(b, s, c, 0.5)

    run_repeated(fn, lambda: [res["keep"]])


def test_dense_detect_is_deterministic(det):
    strides = [8, 16, 32]
    wh = [[[10, 13], [16, 30], [33, 23]], [[30, 61], [62, 45], [59, 119]], [[116, 90], [156, 198], [373, 326]]]
    dh = det.DenseAnchorHead(strides, wh, 80)
    g = gen(5)
    n = 6
    heads = [torch.randn(n, 255, 256 // s, 256 // s, generator=g) for s in strides]
    for h in heads:
        h.view(n, 3, 85, h.shape[2], h.shape[3])[:, :, 4] -= 3.0
    heads = [h.cuda() for h in heads]
    for gate in (False, True):
        r = dh.detect_thresholded(heads, 0.1, 0.5, max_det=300, cand_cap=2048, gate=gate, check=False)
        run_repeated(lambda: dh.detect_thresholded(heads, 0.1, 0.5, max_det=300, cand_cap=2048, gate=gate, check=False, out=r),
                     lambda: [r["idx"], r["count"], r["scores"], r["classes"]])


def test_training_side_is_deterministic(det):
    strides = (4, 8, 16, 32, 64)
    rpn = det.RegionProposalNetwork(list(strides))
    hw = [(256 // s, 256 // s) for s in strides]
    dev = torch.device("cuda")
    anchors = torch.cat(rpn.anchor_generator.grid_anchors(hw, dev), 0)
    grid = rpn.anchor_generator.grid_layout(hw)
    g = gen(7)
    gts = [rand_boxes(int(k), 256.0, g).cuda() for k in (3, 9, 0, 16, 40, 1, 7, 12)]
    n = len(gts)
    table, off = rpn.anchor_matcher.pack_gt(gts, dev)
    obj = [torch.randn(n, 3, h, w, generator=g).cuda() for h, w in hw]
    dlt = [(torch.randn(n, 12, h, w, generator=g) * 0.5).cuda() for h, w in hw]
    st = {}

    def dense():
        st["d"] = rpn.assign(anchors, gts, sample=True, seed=11, grid=grid)

    run_repeated(dense, lambda: [st["d"].labels, st["d"].matched, st["d"].sample_count,
                                 torch.sort(st["d"].samples, dim=1)[0]])

    def generic():
        st["g"] = rpn.assign(anchors, gts, sample=True, seed=11)

    run_repeated(generic, lambda: [st["g"].labels, st["g"].matched])

    def lazy():
        st["l"] = rpn.assign_sampled(anchors, table, off, n, grid, seed=11)
        g_obj, g_dlt = [torch.zeros_like(o) for o in obj], [torch.zeros_like(d) for d in dlt]
        st["sums"] = rpn._run_sampled(anchors, obj, dlt, st["l"], n, None, (g_obj, g_dlt), None)
        st["grads"] = g_obj + g_dlt

    def lazy_snapshot():
        # list order depends on atomic arrival order; the SET of (sample, gt) pairs and the gradients do not
        a = st["l"]
        key = a.samples.long() * 4096 + a.sample_gt.long()
        cnt = a.sample_count
        mask = torch.arange(key.shape[1], device=dev)[None, :] < cnt[:, None]
        key = torch.where(mask, key, torch.full_like(key, -1))
        return [torch.sort(key, dim=1)[0], cnt] + st["grads"]

    run_repeated(lazy, lazy_snapshot)
    # the summed losses are fp32 atomics over CTAs: equal up to summation order
    ref = st["sums"].clone()
    for _ in range(5):
        lazy()
        torch.testing.assert_close(st["sums"], ref, rtol=1e-5, atol=1e-7)


def test_peer_exchange_single_rank_is_deterministic(det):
    dev = torch.device("cuda")
    ps = det.dist.PeerSums(dev)
    vec = torch.arange(8, dtype=torch.float32, device=dev) * 1.5 + 0.25
    assert ps.exchange(vec) is None
    for step in range(200):
        got = ps.exchange(vec + step + 1)
        assert torch.equal(got, vec + step)
    ps.flush()
    ps.check()


def test_peer_exchange_graph_safe_replays(det):
    """det_peer_sums_exchange_dev (step stamp on the device): eager calls and CUDA-graph replays advance the same counter;
    every step returns the previous step's sum; flush() collects the last one (single rank: sum == own vector)."""
    dev = torch.device("cuda")
    ps = det.dist.PeerSums(dev, graph_safe=True)
    vec = torch.arange(8, dtype=torch.float32, device=dev) + 0.5
    out = ps.exchange(vec)            # step 1: nothing to collect yet
    vec += 1.0
    out = ps.exchange(vec)            # step 2: returns step 1's vector
    assert torch.equal(out, torch.arange(8, dtype=torch.float32, device=dev) + 0.5)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        ps.exchange(vec)              # step 3 (warm-up outside capture)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        vec.add_(1.0)
        res = ps.exchange(vec)
    for rep in range(20):             # steps 4..23
        graph.replay()
        torch.cuda.synchronize()
        want = torch.arange(8, dtype=torch.float32, device=dev) + 0.5 + 1.0 + rep  # the vector published one step earlier
        assert torch.equal(res, want), rep
    assert torch.equal(ps.flush(), vec)
    ps.check()
