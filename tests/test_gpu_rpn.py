"""GPU parity: RPN decode from the conv layout, proposal selection (find_top_rpn_proposals) vs the oracle."""
import pytest
import torch

from tests.util import gen, rand_boxes, assert_boxes_close

pytestmark = pytest.mark.gpu

STRIDES = [4, 8, 16, 32, 64]
SIZES = [[32], [64], [128], [256], [512]]
RATIOS = [[0.5, 1.0, 2.0]]


@pytest.fixture(scope="module")
def det():
    import det_b200
    return det_b200


@pytest.fixture(scope="module")
def O():
    from oracle import ref_torch
    return ref_torch


def _heads(n, img, g, scale=1.0):
    obj = [torch.randn(n, 3, img // s, img // s, generator=g) for s in STRIDES]
    dlt = [torch.randn(n, 12, img // s, img // s, generator=g) * scale for s in STRIDES]
    return obj, dlt


def _oracle_decode(O, obj, dlt, img, weights=(1.0, 1.0, 1.0, 1.0)):
    cells = [O.cell_anchors(s, RATIOS[0]) for s in SIZES]
    anchors = O.grid_anchors([(img // s, img // s) for s in STRIDES], STRIDES, cells, 0.0)
    lg, dl = zip(*[O.head_to_hwa(o, d) for o, d in zip(obj, dlt)])
    props = O.decode_proposals(anchors, list(dl), weights)
    return anchors, list(lg), list(dl), props


@pytest.mark.parametrize("img", [448, 224])
def test_decode_heads_matches_oracle(det, O, img):
    g = gen(img)
    obj, dlt = _heads(3, img, g, 0.5)
    rpn = det.RegionProposalNetwork(STRIDES, SIZES, RATIOS)
    logits, boxes, sizes = rpn.decode_heads([o.cuda() for o in obj], [d.cuda() for d in dlt])
    anchors, lg, dl, props = _oracle_decode(O, obj, dlt, img)
    assert sizes == [a.shape[0] for a in anchors]
    if img == 448:
        assert sum(sizes) == 50127
    assert torch.equal(logits.cpu(), torch.cat(lg, 1))          # pure layout change: bit-exact
    assert_boxes_close(boxes.cpu(), torch.cat(props, 1), rtol=1e-5)


def test_reference_signature_decode_proposals(det, O):
    g = gen(1)
    obj, dlt = _heads(2, 128, g)
    anchors, lg, dl, props = _oracle_decode(O, obj, dlt, 128, (10.0, 10.0, 5.0, 5.0))
    rpn = det.RegionProposalNetwork(STRIDES, SIZES, RATIOS, box2box_weights=(10.0, 10.0, 5.0, 5.0))
    got = rpn._decode_proposals([det.Boxes(a.cuda()) for a in anchors], [d.cuda() for d in dl])
    for a, b in zip(got, props):
        assert_boxes_close(a.cpu(), b, rtol=1e-5)


@pytest.mark.parametrize("img,pre,post,training", [(448, 12000, 2000, False), (448, 6000, 1000, True),
                                                   (128, 12000, 2000, False), (64, 50, 20, False)])
def test_find_top_rpn_proposals_matches_oracle(det, O, img, pre, post, training):
    """Kept boxes/logits and counts bit-exact on identical inputs (the proposals fed to both sides are the same)."""
    g = gen(img + pre)
    n = 3
    obj, dlt = _heads(n, img, g, 0.3)
    anchors, lg, dl, props = _oracle_decode(O, obj, dlt, img)
    sizes = [(img, img), (img - 17, img), (img, img - 40)]
    want = O.find_top_rpn_proposals(props, lg, sizes, 0.7, pre, post, 0.0, training)
    got = det.find_top_rpn_proposals([p.cuda() for p in props], [l.cuda() for l in lg], sizes, 0.7, pre, post, 0.0,
                                     training)
    assert len(got) == n
    for i in range(n):
        wb, ws = want[i]
        assert len(got[i]) == wb.shape[0], (i, len(got[i]), wb.shape[0])
        assert torch.equal(got[i].objectness_logits.cpu(), ws)
        assert torch.equal(got[i].proposal_boxes.tensor.cpu(), wb)
        assert got[i].image_size == sizes[i]


def test_find_top_rpn_proposals_min_size_and_nonfinite(det, O):
    g = gen(77)
    obj, dlt = _heads(2, 128, g, 0.3)
    anchors, lg, dl, props = _oracle_decode(O, obj, dlt, 128)
    props[0][0, 5, 2] = float("inf")
    lg[1][1, 7] = float("nan")
    sizes = [(128, 128), (100, 128)]
    want = O.find_top_rpn_proposals([p.clone() for p in props], lg, sizes, 0.7, 1000, 300, 8.0, False)
    got = det.find_top_rpn_proposals([p.cuda() for p in props], [l.cuda() for l in lg], sizes, 0.7, 1000, 300, 8.0, False)
    for i in range(2):
        assert torch.equal(got[i].proposal_boxes.tensor.cpu(), want[i][0])
        assert torch.equal(got[i].objectness_logits.cpu(), want[i][1])
    with pytest.raises(FloatingPointError):
        det.find_top_rpn_proposals([p.cuda() for p in props], [l.cuda() for l in lg], sizes, 0.7, 1000, 300, 8.0, True)


def test_forward_eval_from_heads(det, O):
    g = gen(9)
    obj, dlt = _heads(2, 224, g, 0.3)
    rpn = det.RegionProposalNetwork(STRIDES, SIZES, RATIOS).eval()
    sizes = [(224, 224), (200, 210)]
    props, losses = rpn.forward(sizes, head_outputs=([o.cuda() for o in obj], [d.cuda() for d in dlt]))
    assert losses == {}
    # oracle fed with the GPU's own decoded proposals (identical inputs to the selection stage)
    logits, boxes, lsz = rpn.decode_heads([o.cuda() for o in obj], [d.cuda() for d in dlt])
    pl = list(torch.split(boxes.cpu(), lsz, dim=1))
    ll = list(torch.split(logits.cpu(), lsz, dim=1))
    want = O.find_top_rpn_proposals(pl, ll, sizes, 0.7, 12000, 2000, 0.0, False)
    for i in range(2):
        assert torch.equal(props[i].proposal_boxes.tensor.cpu(), want[i][0])
        assert torch.equal(props[i].objectness_logits.cpu(), want[i][1])


@pytest.mark.parametrize("nms_thresh,pre,post", [(0.7, 2000, 300), (0.05, 2000, 300), (0.3, 4000, 1000), (0.7, 1000, 50)])
def test_find_top_rpn_proposals_tier_cut(det, O, nms_thresh, pre, post):
    """post_nms_topk << candidates: the per-level sweeps first run above a global score cut (exact); with a low NMS
    threshold the cut falls short and the full segments are swept in a second pass."""
    g = gen(int(nms_thresh * 100) + post)
    n = 3
    obj, dlt = _heads(n, 448, g, 0.3)
    anchors, lg, dl, props = _oracle_decode(O, obj, dlt, 448)
    sizes = [(448, 448), (400, 448), (448, 300)]
    want = O.find_top_rpn_proposals(props, lg, sizes, nms_thresh, pre, post, 0.0, False)
    got = det.find_top_rpn_proposals([p.cuda() for p in props], [l.cuda() for l in lg], sizes, nms_thresh, pre, post,
                                     0.0, False)
    for i in range(n):
        wb, ws = want[i]
        assert len(got[i]) == wb.shape[0], (i, len(got[i]), wb.shape[0])
        assert torch.equal(got[i].objectness_logits.cpu(), ws)
        assert torch.equal(got[i].proposal_boxes.tensor.cpu(), wb)


def test_find_top_rpn_proposals_equal_logits_straddle_the_topk_cut(det, O):
    """Quantised logits: the per-level radix select has to go on into the index bits (stable order, lower index first)."""
    g = gen(123)
    n = 2
    obj, dlt = _heads(n, 224, g, 0.3)
    obj = [(o * 2).round() / 2 for o in obj]
    anchors, lg, dl, props = _oracle_decode(O, obj, dlt, 224)
    sizes = [(224, 224), (200, 224)]
    for pre, post in ((100, 40), (700, 300), (3000, 1000)):
        want = O.find_top_rpn_proposals(props, lg, sizes, 0.7, pre, post, 0.0, False)
        got = det.find_top_rpn_proposals([p.cuda() for p in props], [l.cuda() for l in lg], sizes, 0.7, pre, post, 0.0,
                                         False)
        for i in range(n):
            assert torch.equal(got[i].objectness_logits.cpu(), want[i][1])
            assert torch.equal(got[i].proposal_boxes.tensor.cpu(), want[i][0])


@pytest.mark.parametrize("seed", list(range(int(__import__("os").environ.get("DET_STRESS_SEEDS", "12")))))
def test_find_top_rpn_proposals_stress(det, O, seed):
    """Random proposal-selection problems: pyramid size, delta scale (tiny to image-sized boxes), pre/post top-k around the
    level sizes, NMS threshold, min box size, ragged image sizes, quantised logits (ties at the cuts) and an occasional
    non-finite proposal in eval mode: kept boxes, logits and counts bit-exact on identical inputs."""
    g = gen(13000 + seed)
    ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    rf = lambda lo, hi: float(torch.rand(1, generator=g)) * (hi - lo) + lo
    img = 64 * ri(1, 5)
    n = ri(1, 4)
    obj, dlt = _heads(n, img, g, rf(0.05, 1.5))
    if ri(0, 2) == 0:
        obj = [(o * 2).round() / 2 for o in obj]
    anchors, lg, dl, props = _oracle_decode(O, obj, dlt, img)
    if ri(0, 3) == 0:
        props[ri(0, len(props) - 1)][0, 0, ri(0, 3)] = float("inf")
    pre = [5, 60, 1000, 12000][ri(0, 3)]
    post = [1, 20, 300, 2000][ri(0, 3)]
    thr = rf(0.3, 0.9)
    min_size = [0.0, 0.0, 4.0, 20.0][ri(0, 3)]
    sizes = [(img - ri(0, 30), img - ri(0, 30)) for _ in range(n)]
    want = O.find_top_rpn_proposals([p.clone() for p in props], lg, sizes, thr, pre, post, min_size, False)
    got = det.find_top_rpn_proposals([p.cuda() for p in props], [l.cuda() for l in lg], sizes, thr, pre, post, min_size, False)
    for i in range(n):
        wb, ws = want[i]
        assert len(got[i]) == wb.shape[0], (i, len(got[i]), wb.shape[0])
        assert torch.equal(got[i].objectness_logits.cpu(), ws), i
        assert torch.equal(got[i].proposal_boxes.tensor.cpu(), wb), i


_ALT_PATH_SCRIPT = r"""
import sys, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/object-detection-pytorch-rust_b200")
import det_b200 as det
g = torch.Generator().manual_seed(77)
strides = [4, 8, 16, 32, 64]
rpn = det.RegionProposalNetwork(strides)
n = 3
obj = [torch.randn(n, 3, 256 // s, 256 // s, generator=g).cuda() for s in strides]
obj[0] = (obj[0] * 4).round() / 4  # ties at the cuts of the finest level
dlt = [(torch.randn(n, 12, 256 // s, 256 // s, generator=g) * 0.4).cuda() for s in strides]
sizes = torch.tensor([[256, 256], [250, 240], [256, 200]], dtype=torch.int32, device="cuda")
logits, boxes, level_sizes = rpn.decode_heads(obj, dlt)
out = {}
for pre, post in ((2000, 1000), (300, 50), (12000, 2000)):
    b, s, c, f = det.rpn_proposals_batched(boxes, logits, level_sizes, sizes, 0.7, pre, post, 0.0)
    for i in range(n):
        k = int(c[i])
        out[(pre, post, i)] = (b[i, :k].cpu(), s[i, :k].cpu())
torch.save(out, sys.argv[2])
"""


def test_proposal_path_variants_agree_bitwise(tmp_path):
    """The default proposal path (one select CTA per (image, level), per-level tile sort, finish kernel) against the forms
    it replaced and still falls back to for large pyramids (DET_RPN_OLD_SORTS=1: tile sort + merge passes, rekey / pad /
    sort / emit; DET_RPN_SELECT_SPLIT=0: one select CTA per image): same proposals, same logits, same counts, bit for bit.
    The switches are read once per process, so each variant runs in its own interpreter."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "alt_path.py"
    script.write_text(_ALT_PATH_SCRIPT)
    results = []
    for extra in ({}, {"DET_RPN_OLD_SORTS": "1"}, {"DET_RPN_SELECT_SPLIT": "0"}):
        out = tmp_path / ("out_" + "_".join(extra) + ".pt")
        env = dict(os.environ, **extra)
        subprocess.run([sys.executable, str(script), root, str(out)], check=True, env=env, timeout=300)
        results.append(torch.load(out))
    base = results[0]
    assert len(base) == 9 and all(v[0].shape[0] > 0 for v in base.values())
    for other in results[1:]:
        assert other.keys() == base.keys()
        for k in base:
            assert torch.equal(base[k][0], other[k][0]) and torch.equal(base[k][1], other[k][1]), k
