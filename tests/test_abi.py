"""CPU: libdet_b200.so loads and exports exactly the symbols include/det_b200.h declares (no compute calls)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "det_b200.h")).read()
    return re.findall(r"^DET_API\s+[\w\s\*]+?\b(det_[a-z0-9_]+)\s*\(", text, flags=re.M)


def test_header_symbols_are_exported_and_bound():
    from det_b200 import _native as N
    names = _declared()
    assert len(names) >= 20 and len(set(names)) == len(names)
    assert N.missing_symbols() == []
    assert sorted(names) == sorted(N.PROTOTYPES), "ctypes prototypes out of sync with include/det_b200.h"
    for n in names:
        assert N.fn(n) is not None


def test_argument_counts_match_header():
    from det_b200 import _native as N
    text = open(os.path.join(ROOT, "include", "det_b200.h")).read()
    for name, body in re.findall(r"^DET_API\s+[\w\s\*]+?\b(det_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.M | re.S):
        body = body.strip()
        n_args = 0 if body in ("", "void") else body.count(",") + 1
        assert n_args == len(N.PROTOTYPES[name][1]), name


def test_abi_version_and_pure_queries():
    from det_b200 import _native as N
    assert N.fn("det_abi_version")() == 3
    assert N.fn("det_nms_workspace_bytes")(4, 1000) == 256
    assert N.fn("det_nms_workspace_bytes")(2, 25200) > 2 * 25200 * 40
    assert N.fn("det_rpn_proposals_workspace_bytes")(1, 50127) > 50127 * 40
    assert N.fn("det_match_workspace_bytes")(8, 50127, 100) >= 400


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "det_b200.h")).read()
    assert "torch" not in text.lower().replace("pytorch-rust", "").replace("python/src", "").replace("torchvision", "") \
        or "at::" not in text
    assert "at::Tensor" not in text and "#include <torch" not in text


def test_dense_detect_queries_and_argument_checks_need_no_gpu():
    """Pure size query + argument validation of det_dense_detect: rejected before any CUDA call is made."""
    import ctypes
    from det_b200 import _native as N
    wsb = N.fn("det_dense_detect_workspace_bytes")
    assert wsb(0, 1024) == 256
    assert wsb(32, 2048) == 32 * 128 + 32 * 2048 * 28          # padded counters + (box, score, class, row) lists
    assert N.fn("det_dense_detect_counter_bytes")(32) == 32 * 128 and N.fn("det_dense_detect_counter_bytes")(0) == 0
    assert N.fn("det_nms_workspace_bytes")(32, 25200) > wsb(32, 4096)   # the large path carries the top-k tier's lists
    f = N.fn("det_dense_detect")
    null = ctypes.c_void_p(0)
    args = lambda cap, max_det, mode: (null, 0, 4, 3, 80, 4.0, 0.1, 0.5, mode, 1, cap, max_det, null, null, null, null,
                                       null, null, null, 0, null)
    assert f(*args(5000, 300, 0)) == -1     # cand_cap > 4096
    assert f(*args(1024, 0, 0)) == -1       # max_det < 1
    assert f(*args(1024, 300, 7)) == -1     # unknown NMS mode
    assert f(*args(1024, 300, 0)) == -1     # null outputs
    assert b"null output" in N.fn("det_last_error")()


def test_peer_sums_argument_checks_need_no_gpu():
    import ctypes
    from det_b200 import _native as N
    one = ctypes.c_void_p(16)  # never dereferenced: every call below is rejected by the argument checks
    pub, col, exc = N.fn("det_peer_sums_publish"), N.fn("det_peer_sums_collect"), N.fn("det_peer_sums_exchange")
    assert pub(one, 13, 0, 2, one, 8, 0, 1, None) == -1          # width > 12
    assert pub(one, 8, 2, 2, one, 8, 0, 1, None) == -1           # rank >= world
    assert pub(one, 8, 0, 2, one, 8, 8, 1, None) == -1           # slot >= slots
    assert col(one, 8, 33, one, 8, 0, 1, 10 ** 9, None, None) == -1   # world > 32
    assert col(one, 8, 2, one, 8, 0, 1, 0, None, None) == -1          # timeout must be positive
    assert exc(one, one, 8, 0, 2, one, 3, 1, 1, 10 ** 9, None, None) == -1   # slots < 4
    assert exc(one, one, 8, 0, 2, one, 8, 1, 6, 10 ** 9, None, None) == -1   # lag > slots - 3
